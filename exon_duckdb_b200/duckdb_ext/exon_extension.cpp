// exon_extension.cpp -- the DuckDB-facing host of the B200 scan engine (`LOAD exon`).
//
// Same surface as the reference extension for the read_fasta / read_fastq path:
//   exon_init / exon_version                       <- exon/src/exon_extension.cpp:109-121
//   read_fasta(VARCHAR, compression := VARCHAR)    <- arrow_table_function/module.cpp:296-318 (Register)
//   read_fastq(VARCHAR, compression := VARCHAR)       schema as FileTypeBind produces it (:75-156)
//   FROM 'x.fasta[.gz]' / 'x.fastq[.gz]'           <- ReplacementScan (:320-382)
//   gc_content, reverse_complement, complement,
//   transcribe, reverse_transcribe, translate_dna_to_aa  <- sequence_functions/module.cpp:30-370
//   quality_score_string_to_list                   <- fastq_functions/module.cpp:28-54
//   COPY ... TO (FORMAT 'fastq' | 'fasta')         <- fastq_functions/module.hpp:30 (declared, removed from the reference)
// It is not the reference's glue: there is no Arrow stream between the engine and DuckDB.  bind / init_global /
// scan call the C ABI of libexon_b200.so (include/exon_b200.h, exb_reader_*) and fill the DataChunk directly:
// string_t values borrow the reader's host buffers, which a VectorBuffer attached to the vector keeps alive.
//   * projection push-down: only the requested columns are gathered on the device and copied back;
//   * COUNT(*) (row-id only): rows are counted on the device, nothing is materialised (exb_reader_count);
//   * simple filters (TableFilterSet) are serialised exactly like the reference's FilterToString (:158-214) and
//     evaluated on the device;
//   * complex filters  list_avg(quality_score_string_to_list(quality_scores)) <op> c  and
//     gc_content(sequence) <op> c  are absorbed through pushdown_complex_filter and become per-record
//     predicates evaluated by the scan kernels' outputs (mean_quality(...) / gc_content(...) terms).
// The scalar functions run the CUDA kernels on each DataChunk through exb_*_host.  There is no CPU fallback:
// without a CUDA device every one of these functions raises an error.
#define DUCKDB_EXTENSION_MAIN

#include <cmath>
#include <cstring>
#include <mutex>
#include <thread>

#include "duckdb.hpp"
#include "duckdb/common/enums/expression_type.hpp"
#include "duckdb/common/types/vector_buffer.hpp"
#include "duckdb/common/file_system.hpp"
#include "duckdb/function/copy_function.hpp"
#include "duckdb/function/replacement_scan.hpp"
#include "duckdb/parser/parsed_data/create_copy_function_info.hpp"
#include "duckdb/function/scalar_function.hpp"
#include "duckdb/function/table_function.hpp"
#include "duckdb/main/extension_util.hpp"
#include "duckdb/parser/expression/constant_expression.hpp"
#include "duckdb/parser/expression/function_expression.hpp"
#include "duckdb/parser/parsed_data/create_scalar_function_info.hpp"
#include "duckdb/parser/parsed_data/create_table_function_info.hpp"
#include "duckdb/parser/tableref/table_function_ref.hpp"
#include "duckdb/planner/expression/bound_cast_expression.hpp"
#include "duckdb/planner/expression/bound_columnref_expression.hpp"
#include "duckdb/planner/expression/bound_comparison_expression.hpp"
#include "duckdb/planner/expression/bound_constant_expression.hpp"
#include "duckdb/planner/expression/bound_function_expression.hpp"
#include "duckdb/planner/filter/conjunction_filter.hpp"
#include "duckdb/planner/filter/constant_filter.hpp"
#include "duckdb/optimizer/optimizer_extension.hpp"
#include "duckdb/planner/expression_iterator.hpp"
#include "duckdb/planner/logical_operator_visitor.hpp"
#include "duckdb/planner/operator/logical_filter.hpp"
#include "duckdb/planner/operator/logical_get.hpp"
#include "duckdb/planner/table_filter.hpp"

#include "../../include/exon_b200.h"

namespace exon_b200 {
using namespace duckdb;

// ------------------------------------------------------------------ table function
struct ScanInfo : public TableFunctionInfo {
	explicit ScanInfo(string file_type_p) : file_type(std::move(file_type_p)) {
	}
	string file_type;
};

// a scalar function of the reference applied to a column of the scan, served by the scan itself (exb_computed)
struct ComputedColumn {
	int32_t kind; // EXB_C_*
	int32_t arg;  // EXB_C_SEQ_MAP: EXB_MAP_*
	string name;  // how EXPLAIN shows it
	LogicalType type;
};

struct ScanBindData : public TableFunctionData {
	string file_name;
	string file_type;
	string compression; // "auto_detect" = infer from the extension (module.cpp:85,89-101)
	int32_t gpus = 0;   // named parameter `gpus`: 0 = all visible devices when the input is large enough
	vector<string> names;
	vector<string> pushed;           // predicates absorbed by pushdown_complex_filter, in the engine's filter syntax
	vector<ComputedColumn> computed; // projections absorbed by the optimizer extension: column id names.size() + k
	vector<bool> dead;               // file columns nothing above the scan reads any more (after the rewrite)
};

// keeps the reader's host buffers alive for as long as a vector points into them
struct BatchHolder {
	explicit BatchHolder(exb_batch batch_p) : batch(batch_p) {
	}
	~BatchHolder() {
		exb_batch_release(&batch);
	}
	exb_batch batch;
};
class BatchBuffer : public VectorBuffer { // VARCHAR vectors: joins the vector's string heap references
public:
	explicit BatchBuffer(shared_ptr<BatchHolder> holder_p) : VectorBuffer(VectorBufferType::OPAQUE_BUFFER), holder(std::move(holder_p)) {
	}
	shared_ptr<BatchHolder> holder;
};
struct BatchAux : public VectorAuxiliaryData { // fixed-width vectors, like ArrowAuxiliaryData (arrow_aux_data.hpp:16-25)
	explicit BatchAux(shared_ptr<BatchHolder> holder_p) : VectorAuxiliaryData(VectorAuxiliaryDataType::ARROW_AUXILIARY), holder(std::move(holder_p)) {
	}
	shared_ptr<BatchHolder> holder;
};

// what ArrowScanLocalState::batch_index is to the reference (arrow.hpp, ArrowGetBatchIndex): the position of the chunk a
// thread is holding in the order of the file, so that order-preserving sinks (batch collector, LIMIT, COPY) can run
struct ScanLocalState : public LocalTableFunctionState {
	idx_t batch_index = 0;
	int64_t shard = -1; // the reader this thread pulls from
};

// The reference streams a file through ONE thread (ArrowScanGlobalState::max_threads = 1, arrow.hpp:107-119).  Here a
// scan owns one device pipeline (reader) per GPU: readers[k] parses the k-th byte range of the file (or the k-th group
// of files of a directory) on device k, and every DuckDB scan thread pulls finished 2048-row batches from one of them.
// Batches are built on the device in DuckDB's own vector layout, so a scan call assigns pointers and returns.
struct ScanGlobalState : public GlobalTableFunctionState {
	std::mutex lock;
	vector<exb_reader *> readers;
	idx_t threads_per_reader = 1;
	idx_t next_slot = 0; // slot s pulls from readers[s / threads_per_reader]; slots are claimed in increasing order
	vector<column_t> column_ids;
	idx_t n_file_cols = 0;
	vector<bool> dead;
	bool count_only = false;
	int64_t count_left = 0;
	bool done = false;
	~ScanGlobalState() override {
		for (auto r : readers) {
			exb_reader_close(r);
		}
	}
	idx_t MaxThreads() const override {
		return count_only ? 1 : readers.size() * threads_per_reader;
	}
};

static unique_ptr<FunctionData> ScanBind(ClientContext &context, TableFunctionBindInput &input, vector<LogicalType> &return_types,
                                         vector<string> &names) {
	auto &info = input.info->Cast<ScanInfo>();
	auto result = make_uniq<ScanBindData>();
	result->file_name = input.inputs[0].GetValue<string>();
	result->file_type = info.file_type;
	result->compression = "auto_detect";
	for (auto &kv : input.named_parameters) {
		if (kv.first == "compression") {
			result->compression = kv.second.GetValue<string>();
		} else if (kv.first == "gpus") {
			result->gpus = kv.second.GetValue<int32_t>();
			if (result->gpus < 0) {
				throw InvalidInputException("read_%s: gpus must be >= 0", info.file_type.c_str());
			}
		}
	}
	// the reference opens a reader at bind time to learn the schema, so a bad path / codec fails here (module.cpp:95-108)
	exb_reader *probe = nullptr;
	const char *comp = result->compression == "auto_detect" ? nullptr : result->compression.c_str();
	if (exb_reader_open(result->file_name.c_str(), info.file_type.c_str(), comp, STANDARD_VECTOR_SIZE, nullptr, 0, &probe) != 0) {
		throw std::runtime_error(exb_last_error());
	}
	const char *cols[4];
	int n = exb_reader_columns(probe, cols, 4);
	for (int c = 0; c < n; c++) {
		names.emplace_back(cols[c]);
		return_types.emplace_back(LogicalType::VARCHAR);
	}
	exb_reader_close(probe);
	result->names = names;
	result->dead.assign(names.size(), false);
	return std::move(result);
}

// FilterToString of the reference (module.cpp:158-214): same text, so both hosts speak one filter syntax
static string FilterToString(const TableFilter &filter, const string &column_name) {
	switch (filter.filter_type) {
	case TableFilterType::CONSTANT_COMPARISON: {
		auto &f = (const ConstantFilter &)filter;
		return column_name + ExpressionTypeToOperator(f.comparison_type) + f.constant.ToSQLString();
	}
	case TableFilterType::CONJUNCTION_AND: {
		auto &f = (const ConjunctionAndFilter &)filter;
		vector<string> parts;
		for (auto &child : f.child_filters) {
			parts.push_back(FilterToString(*child, column_name));
		}
		return "(" + StringUtil::Join(parts, " AND ") + ")";
	}
	case TableFilterType::CONJUNCTION_OR: {
		auto &f = (const ConjunctionOrFilter &)filter;
		vector<string> parts;
		for (auto &child : f.child_filters) {
			parts.push_back(FilterToString(*child, column_name));
		}
		return "(" + StringUtil::Join(parts, " OR ") + ")";
	}
	case TableFilterType::IS_NOT_NULL:
		return column_name + " IS NOT NULL";
	case TableFilterType::IS_NULL:
		return column_name + " IS NULL";
	default:
		throw NotImplementedException("FilterToString: filter type not implemented");
	}
}

// How many device pipelines a scan gets (see the policy below).  Byte-range shards need ONE uncompressed
// file; a directory is split into groups of whole files of about equal size.
static idx_t PlanReaders(const ScanBindData &bind, const char *comp, int64_t &total_bytes, int32_t &n_files, int32_t &shardable) {
	total_bytes = 0;
	n_files = 1;
	shardable = 0;
	if (exb_reader_plan(bind.file_name.c_str(), bind.file_type.c_str(), comp, &total_bytes, &n_files, &shardable) != 0) {
		throw std::runtime_error(exb_last_error());
	}
	const int devices = exb_device_count();
	// `gpus = n` asks for n device pipelines.  Without it the scan uses ONE unless EXON_B200_GPUS says otherwise ("all" or a
	// number; then every device gets a pipeline as long as each is left with >= 256 MiB of input).  One is the default
	// because a pipeline is fed by the host's page-cache copy workers, which several pipelines share: measured on a 2-GPU box
	// (profiles/round2_duckdb_multigpu.txt) two pipelines help queries that ship little back (AVG(gc_content): 1.35 x) and
	// hurt the ones that materialise columns for order-preserving sinks.
	int env_gpus = 1;
	if (const char *e = getenv("EXON_B200_GPUS")) {
		env_gpus = StringUtil::Lower(e) == "all" ? devices : atoi(e);
	}
	idx_t want = bind.gpus > 0 ? (idx_t)bind.gpus : (idx_t)MaxValue<int>(MinValue<int>(env_gpus, devices), 1);
	if (bind.gpus > 0 && bind.gpus > devices) {
		throw InvalidInputException("gpus := %d, but only %d CUDA device(s) are visible", bind.gpus, devices);
	}
	if (bind.gpus == 0) {
		want = MinValue<idx_t>(want, (idx_t)MaxValue<int64_t>(1, total_bytes / (256ll << 20)));
	}
	if (const char *force = getenv("EXON_B200_FORCE_READERS")) { // tests: several pipelines on however many devices there are
		if (atoi(force) > 0) {
			want = (idx_t)atoi(force);
		}
	}
	if (shardable) {
		return MaxValue<idx_t>(want, 1);
	}
	if (n_files > 1) {
		return MaxValue<idx_t>(MinValue<idx_t>(want, (idx_t)n_files), 1);
	}
	return 1;
}

static unique_ptr<GlobalTableFunctionState> ScanInitGlobal(ClientContext &context, TableFunctionInitInput &input) {
	auto &bind = input.bind_data->Cast<ScanBindData>();
	auto state = make_uniq<ScanGlobalState>();
	state->column_ids = input.column_ids;
	state->n_file_cols = bind.names.size();
	state->dead = bind.dead;

	vector<string> terms;
	if (input.filters) {
		for (auto &f : input.filters->filters) {
			auto col = input.column_ids[f.first]; // filter keys index column_ids, not table columns
			terms.push_back(FilterToString(*f.second, bind.names[col]));
		}
	}
	for (auto &p : bind.pushed) {
		terms.push_back(p);
	}
	const string filter_clause = StringUtil::Join(terms, " AND ");

	exb_reader_options opt;
	memset(&opt, 0, sizeof(opt));
	opt.size = sizeof(opt);
	opt.flags = EXB_RD_STRING_T | EXB_RD_NO_OFFSETS;
	bool any_column = false;
	vector<bool> computed_used(bind.computed.size(), false);
	for (auto id : input.column_ids) {
		if (id == COLUMN_IDENTIFIER_ROW_ID) {
			continue;
		}
		any_column = true;
		if (id < bind.names.size()) {
			if (!bind.dead[id]) {
				opt.column_mask |= 1u << id;
			}
		} else {
			computed_used[id - bind.names.size()] = true;
		}
	}
	// the reader numbers its computed columns 0, 1, ...: keep the bind data's numbering (unused ones are still computed;
	// the optimizer only creates columns that an expression reads)
	opt.n_computed = (int32_t)bind.computed.size();
	for (idx_t k = 0; k < bind.computed.size(); k++) {
		opt.computed[k].kind = bind.computed[k].kind;
		opt.computed[k].arg = bind.computed[k].arg;
	}

	const char *comp = bind.compression == "auto_detect" ? nullptr : bind.compression.c_str();
	int64_t total_bytes = 0;
	int32_t n_files = 1, shardable = 0;
	const idx_t n_readers = PlanReaders(bind, comp, total_bytes, n_files, shardable);
	for (idx_t k = 0; k < n_readers; k++) {
		exb_reader_options o = opt;
		o.device = n_readers > 1 || bind.gpus > 0 ? (int32_t)(k % (idx_t)MaxValue<int>(exb_device_count(), 1)) : -1;
		if (n_readers > 1 && shardable) {
			o.range_lo = total_bytes * (int64_t)k / (int64_t)n_readers;
			o.range_hi = k + 1 == n_readers ? 0 : total_bytes * (int64_t)(k + 1) / (int64_t)n_readers;
		} else if (n_readers > 1) {
			o.file_lo = (int32_t)((int64_t)n_files * (int64_t)k / (int64_t)n_readers);
			o.file_hi = (int32_t)((int64_t)n_files * (int64_t)(k + 1) / (int64_t)n_readers);
		}
		exb_reader *reader = nullptr;
		if (exb_reader_open2(bind.file_name.c_str(), bind.file_type.c_str(), comp, STANDARD_VECTOR_SIZE, filter_clause.c_str(), &o, &reader) != 0) {
			throw std::runtime_error(exb_last_error());
		}
		state->readers.push_back(reader);
	}
	// one consumer per device pipeline is enough to hand out pointers; a lone pipeline gets a few so that whatever sits
	// above the scan (string functions, aggregates) runs on several cores
	const idx_t cores = MaxValue<idx_t>(std::thread::hardware_concurrency(), 1);
	// consumers in total: up to 8 (half the cores), spread over the pipelines, at least one each
	state->threads_per_reader = MaxValue<idx_t>(1, MaxValue<idx_t>(2, MinValue<idx_t>(8, cores / 2)) / n_readers);

	if (!any_column) { // COUNT(*): arrow_conversion.cpp:813-816 is the reference's "row id only" case
		state->count_only = true;
		vector<int64_t> rows(n_readers, 0);
		vector<string> errors(n_readers);
		vector<std::thread> workers;
		for (idx_t k = 0; k < n_readers; k++) {
			workers.emplace_back([&, k] {
				if (exb_reader_count(state->readers[k], &rows[k]) != 0) {
					errors[k] = exb_last_error();
				}
			});
		}
		for (auto &w : workers) {
			w.join();
		}
		for (idx_t k = 0; k < n_readers; k++) {
			if (!errors[k].empty()) {
				throw std::runtime_error(errors[k]);
			}
			state->count_left += rows[k];
		}
	}
	return std::move(state);
}

// one output vector <- one column view of a device-built batch: pointer assignments, no per-row work
static void FillVector(Vector &vec, const exb_column_view &col, idx_t n, const shared_ptr<BatchHolder> &holder) {
	vec.SetVectorType(VectorType::FLAT_VECTOR);
	auto &validity = FlatVector::Validity(vec);
	if (col.valid && col.chunk_nulls != 0) {
		if (col.valid_bits) {
			validity.Initialize(const_cast<validity_t *>(reinterpret_cast<const validity_t *>(col.valid_bits)));
		} else {
			for (idx_t i = 0; i < n; i++) {
				if (!col.valid[i]) {
					validity.SetInvalid(i);
				}
			}
		}
	}
	switch (col.type) {
	case EXB_T_VARCHAR:
		if (!col.strings) {
			throw InternalException("exon_b200: a column that was projected out is requested");
		}
		// string_t entries built on the device: <= 12 bytes inlined, longer values point into the batch's host buffer
		FlatVector::SetData(vec, const_cast<data_ptr_t>(reinterpret_cast<const_data_ptr_t>(col.strings)));
		StringVector::AddBuffer(vec, make_buffer<BatchBuffer>(holder));
		break;
	case EXB_T_INT32_LIST: {
		FlatVector::SetData(vec, const_cast<data_ptr_t>(reinterpret_cast<const_data_ptr_t>(col.list_entries)));
		auto &child = ListVector::GetEntry(vec);
		child.SetVectorType(VectorType::FLAT_VECTOR);
		FlatVector::SetData(child, const_cast<data_ptr_t>(reinterpret_cast<const_data_ptr_t>(col.values)));
		ListVector::SetListSize(vec, (idx_t)col.n_values);
		child.GetBuffer()->SetAuxiliaryData(make_uniq<BatchAux>(holder));
		vec.GetBuffer()->SetAuxiliaryData(make_uniq<BatchAux>(holder));
		break;
	}
	default:
		FlatVector::SetData(vec, const_cast<data_ptr_t>(reinterpret_cast<const_data_ptr_t>(col.values)));
		vec.GetBuffer()->SetAuxiliaryData(make_uniq<BatchAux>(holder));
		break;
	}
}

static void ScanFunction(ClientContext &context, TableFunctionInput &input, DataChunk &output) {
	auto &state = input.global_state->Cast<ScanGlobalState>();
	auto &local = input.local_state->Cast<ScanLocalState>();
	if (state.count_only) {
		std::lock_guard<std::mutex> guard(state.lock);
		const idx_t n = (idx_t)MinValue<int64_t>(STANDARD_VECTOR_SIZE, state.count_left);
		state.count_left -= n;
		output.SetCardinality(n);
		return;
	}
	exb_batch batch;
	for (;;) {
		if (local.shard < 0) {
			std::lock_guard<std::mutex> guard(state.lock);
			if (state.next_slot >= state.readers.size() * state.threads_per_reader) {
				output.SetCardinality(0);
				return;
			}
			local.shard = (int64_t)(state.next_slot++ / state.threads_per_reader);
		}
		if (exb_reader_next(state.readers[local.shard], &batch) != 0) { // thread safe: a batch goes to exactly one caller
			throw InvalidInputException(exb_last_error());
		}
		if (batch.n_rows > 0) {
			break;
		}
		local.shard = -1; // this pipeline is drained: move on to the next unclaimed one (always a later shard)
	}
	// shard-major order = file order: shard k holds the k-th byte range (or the k-th group of files)
	local.batch_index = ((idx_t)local.shard << 32) | (idx_t)batch.batch_index;
	const idx_t n = (idx_t)batch.n_rows;
	auto holder = make_shared<BatchHolder>(batch);
	for (idx_t out_col = 0; out_col < state.column_ids.size(); out_col++) {
		const auto id = state.column_ids[out_col];
		auto &vec = output.data[out_col];
		if (id == COLUMN_IDENTIFIER_ROW_ID || (id < state.n_file_cols && state.dead[id])) {
			vec.SetVectorType(VectorType::CONSTANT_VECTOR);
			ConstantVector::SetNull(vec, true);
			continue;
		}
		FillVector(vec, batch.cols[id], n, holder);
	}
	output.SetCardinality(n);
}

static string ScanToString(const FunctionData *bind_data_p) {
	auto &bind = bind_data_p->Cast<ScanBindData>();
	string s = bind.file_name;
	if (!bind.pushed.empty()) {
		s += "\nDevice filters: " + StringUtil::Join(bind.pushed, " AND ");
	}
	if (!bind.computed.empty()) {
		vector<string> names;
		for (auto &c : bind.computed) {
			names.push_back(c.name);
		}
		s += "\nDevice columns: " + StringUtil::Join(names, ", ");
	}
	vector<string> dead;
	for (idx_t c = 0; c < bind.dead.size(); c++) {
		if (bind.dead[c]) {
			dead.push_back(bind.names[c]);
		}
	}
	if (!dead.empty()) {
		s += "\nNot read back: " + StringUtil::Join(dead, ", ");
	}
	return s;
}

// table_scan_progress: input bytes the device pipelines have consumed / input bytes (the reference leaves it unset,
// module.cpp:296-318)
static double ScanProgress(ClientContext &context, const FunctionData *bind_data, const GlobalTableFunctionState *global_state) {
	auto &state = global_state->Cast<ScanGlobalState>();
	int64_t done = 0, total = 0;
	for (auto r : state.readers) {
		int64_t d = 0, t = 0;
		exb_reader_progress(r, &d, &t);
		done += d;
		total += t;
	}
	return total > 0 ? 100.0 * (double)done / (double)total : 100.0;
}

// ---- complex filter push-down
static const Expression &StripCast(const Expression &e, bool &was_cast, LogicalTypeId &cast_to) {
	was_cast = false;
	if (e.expression_class == ExpressionClass::BOUND_CAST) {
		auto &c = e.Cast<BoundCastExpression>();
		was_cast = true;
		cast_to = c.return_type.id();
		return *c.child;
	}
	return e;
}

static bool IsColumn(const Expression &e, LogicalGet &get, const ScanBindData &bind, const char *name) {
	if (e.expression_class != ExpressionClass::BOUND_COLUMN_REF) {
		return false;
	}
	auto &ref = e.Cast<BoundColumnRefExpression>();
	if (ref.depth != 0 || ref.binding.table_index != get.table_index || ref.binding.column_index >= get.column_ids.size()) {
		return false;
	}
	auto id = get.column_ids[ref.binding.column_index];
	return id < bind.names.size() && bind.names[id] == name;
}

static bool NumericConstant(const Expression &e, double &value, LogicalTypeId &type) {
	if (e.expression_class != ExpressionClass::BOUND_CONSTANT) {
		return false;
	}
	auto &c = e.Cast<BoundConstantExpression>();
	if (c.value.IsNull()) {
		return false;
	}
	type = c.value.type().id();
	if (type != LogicalTypeId::DOUBLE && type != LogicalTypeId::FLOAT) {
		return false;
	}
	value = c.value.GetValue<double>();
	// DuckDB orders DOUBLE totally (NaN is the largest value, NaN = NaN); the device compares with IEEE semantics, and
	// nobody re-checks an absorbed predicate: leave NaN / infinite constants to DuckDB's own filter
	return std::isfinite(value);
}

static const char *OperatorText(ExpressionType t, bool flipped) {
	switch (t) {
	case ExpressionType::COMPARE_GREATERTHAN:
		return flipped ? "<" : ">";
	case ExpressionType::COMPARE_GREATERTHANOREQUALTO:
		return flipped ? "<=" : ">=";
	case ExpressionType::COMPARE_LESSTHAN:
		return flipped ? ">" : "<";
	case ExpressionType::COMPARE_LESSTHANOREQUALTO:
		return flipped ? ">=" : "<=";
	case ExpressionType::COMPARE_EQUAL:
		return "=";
	case ExpressionType::COMPARE_NOTEQUAL:
		return "!=";
	default:
		return nullptr;
	}
}

// list_avg(quality_score_string_to_list(quality_scores)): list_avg(l) is the macro list_aggr(l, 'avg')
// (duckdb/src/catalog/default/default_functions.cpp:103)
static bool IsMeanQuality(const Expression &f, LogicalGet &get, const ScanBindData &bind) {
	if (bind.file_type != "fastq" || f.expression_class != ExpressionClass::BOUND_FUNCTION) {
		return false;
	}
	auto &fn = f.Cast<BoundFunctionExpression>();
	const string fname = StringUtil::Lower(fn.function.name);
	if (!(fname == "list_aggr" || fname == "list_aggregate" || fname == "array_aggr" || fname == "array_aggregate") || fn.children.size() != 2 ||
	    fn.return_type.id() != LogicalTypeId::DOUBLE) {
		return false;
	}
	auto &name_arg = *fn.children[1];
	if (name_arg.expression_class != ExpressionClass::BOUND_CONSTANT) {
		return false;
	}
	auto &cv = name_arg.Cast<BoundConstantExpression>().value;
	if (cv.IsNull() || cv.type().id() != LogicalTypeId::VARCHAR) {
		return false;
	}
	const string agg = StringUtil::Lower(cv.GetValue<string>());
	if (agg != "avg" && agg != "mean") {
		return false;
	}
	auto &inner = *fn.children[0];
	if (inner.expression_class != ExpressionClass::BOUND_FUNCTION) {
		return false;
	}
	auto &qfn = inner.Cast<BoundFunctionExpression>();
	return StringUtil::Lower(qfn.function.name) == "quality_score_string_to_list" && qfn.children.size() == 1 &&
	       IsColumn(*qfn.children[0], get, bind, "quality_scores");
}

// returns the engine predicate for `func_side <op> constant`, or "" if the expression is not one we absorb
static string MatchPredicate(const Expression &func_side, const Expression &const_side, ExpressionType cmp, bool flipped, LogicalGet &get,
                             const ScanBindData &bind) {
	const char *op = OperatorText(cmp, flipped);
	double value;
	LogicalTypeId ctype;
	if (!op || !NumericConstant(const_side, value, ctype)) {
		return "";
	}
	bool was_cast;
	LogicalTypeId cast_to = LogicalTypeId::INVALID;
	const Expression &f = StripCast(func_side, was_cast, cast_to);
	if (f.expression_class != ExpressionClass::BOUND_FUNCTION) {
		return "";
	}
	auto &fn = f.Cast<BoundFunctionExpression>();
	const string fname = StringUtil::Lower(fn.function.name);
	char buf[64];
	snprintf(buf, sizeof(buf), "%.17g", value);
	if (!was_cast && ctype == LogicalTypeId::DOUBLE && IsMeanQuality(f, get, bind)) {
		return string("mean_quality(quality_scores)") + op + buf;
	}
	if (fname == "gc_content" && fn.children.size() == 1 && IsColumn(*fn.children[0], get, bind, "sequence")) {
		// FLOAT result compared as FLOAT (constant cast to FLOAT) or widened to DOUBLE: both are exact in double
		if ((!was_cast && ctype == LogicalTypeId::FLOAT) || (was_cast && cast_to == LogicalTypeId::DOUBLE && ctype == LogicalTypeId::DOUBLE)) {
			return string("gc_content(sequence)") + op + buf;
		}
	}
	return "";
}

static void ScanPushdownComplexFilter(ClientContext &context, LogicalGet &get, FunctionData *bind_data_p,
                                      vector<unique_ptr<Expression>> &filters) {
	auto &bind = bind_data_p->Cast<ScanBindData>();
	for (idx_t i = 0; i < filters.size();) {
		string pred;
		auto &e = *filters[i];
		if (e.expression_class == ExpressionClass::BOUND_COMPARISON) {
			auto &cmp = e.Cast<BoundComparisonExpression>();
			pred = MatchPredicate(*cmp.left, *cmp.right, cmp.type, false, get, bind);
			if (pred.empty()) {
				pred = MatchPredicate(*cmp.right, *cmp.left, cmp.type, true, get, bind);
			}
		}
		if (pred.empty()) {
			i++;
			continue;
		}
		bind.pushed.push_back(pred);
		filters.erase(filters.begin() + i); // the scan applies it; nobody re-checks (pushdown_get.cpp:26-47)
	}
}

static unique_ptr<LocalTableFunctionState> ScanInitLocal(ExecutionContext &context, TableFunctionInitInput &input,
                                                        GlobalTableFunctionState *global_state) {
	return make_uniq<ScanLocalState>();
}

// ArrowScanCardinality (arrow.cpp): unknown -- the row count of a text file is not known before it is read
static unique_ptr<NodeStatistics> ScanCardinality(ClientContext &context, const FunctionData *bind_data) {
	return make_uniq<NodeStatistics>();
}

// ArrowGetBatchIndex (arrow.cpp)
static idx_t ScanGetBatchIndex(ClientContext &context, const FunctionData *bind_data, LocalTableFunctionState *local_state,
                               GlobalTableFunctionState *global_state) {
	return local_state->Cast<ScanLocalState>().batch_index;
}

// ---- projection push-down of the scalar functions (north_star kernel 3)
// An OptimizerExtension (optimizer_extension.hpp; run at the end of Optimizer::Optimize, optimizer.cpp:162-166) that
// rewrites  gc_content(sequence), reverse_complement / complement / transcribe / reverse_transcribe(sequence),
// quality_score_string_to_list(quality_scores) and list_avg(quality_score_string_to_list(quality_scores))
// directly above a read_fasta / read_fastq scan into hidden computed columns OF THE SCAN: the kernels that would run per
// DataChunk through the scalar-function callbacks (sequence_functions/module.cpp:30-253, fastq_functions/module.cpp:
// 32-50) run once per 64 MiB chunk while it is in HBM, and a source column nothing else reads never crosses PCIe.
// Signatures, results and error messages are unchanged; EXPLAIN lists the columns under "Device columns".
struct RewriteCtx {
	LogicalGet &get;
	ScanBindData &bind;
	bool allow_throwing; // the LUT maps raise on a byte outside their table: only where every scanned row is evaluated
	idx_t rewritten = 0;
};

static bool MatchComputed(const Expression &e, RewriteCtx &ctx, ComputedColumn &out) {
	if (e.expression_class != ExpressionClass::BOUND_FUNCTION) {
		return false;
	}
	auto &fn = e.Cast<BoundFunctionExpression>();
	const string fname = StringUtil::Lower(fn.function.name);
	if (IsMeanQuality(e, ctx.get, ctx.bind)) {
		out = ComputedColumn {EXB_C_MEAN_QUALITY, 0, "list_avg(quality_score_string_to_list(quality_scores))", LogicalType::DOUBLE};
		return true;
	}
	if (fn.children.size() != 1) {
		return false;
	}
	if (fname == "gc_content" && fn.return_type.id() == LogicalTypeId::FLOAT && IsColumn(*fn.children[0], ctx.get, ctx.bind, "sequence")) {
		out = ComputedColumn {EXB_C_GC_CONTENT, 0, "gc_content(sequence)", LogicalType::FLOAT};
		return true;
	}
	if (fname == "quality_score_string_to_list" && ctx.bind.file_type == "fastq" && IsColumn(*fn.children[0], ctx.get, ctx.bind, "quality_scores")) {
		out = ComputedColumn {EXB_C_QUALITY_LIST, 0, "quality_score_string_to_list(quality_scores)", LogicalType::LIST(LogicalType::INTEGER)};
		return true;
	}
	static const struct {
		const char *name;
		int mode;
	} maps[] = {{"reverse_complement", EXB_MAP_REVERSE_COMPLEMENT},
	            {"complement", EXB_MAP_COMPLEMENT},
	            {"transcribe", EXB_MAP_TRANSCRIBE},
	            {"reverse_transcribe", EXB_MAP_REVERSE_TRANSCRIBE}};
	for (auto &m : maps) {
		if (fname == m.name && ctx.allow_throwing && fn.return_type.id() == LogicalTypeId::VARCHAR &&
		    IsColumn(*fn.children[0], ctx.get, ctx.bind, "sequence")) {
			out = ComputedColumn {EXB_C_SEQ_MAP, m.mode, string(m.name) + "(sequence)", LogicalType::VARCHAR};
			return true;
		}
	}
	return false;
}

// the scan column (position in get.column_ids) that serves `c`, created on first use
static idx_t ComputedBinding(RewriteCtx &ctx, const ComputedColumn &c) {
	idx_t k = 0;
	for (; k < ctx.bind.computed.size(); k++) {
		if (ctx.bind.computed[k].kind == c.kind && ctx.bind.computed[k].arg == c.arg) {
			break;
		}
	}
	if (k == ctx.bind.computed.size()) {
		if (k >= EXB_MAX_COMPUTED) {
			return DConstants::INVALID_INDEX;
		}
		ctx.bind.computed.push_back(c);
		ctx.get.returned_types.push_back(c.type);
		ctx.get.names.push_back(c.name);
	}
	const column_t id = ctx.bind.names.size() + k;
	for (idx_t i = 0; i < ctx.get.column_ids.size(); i++) {
		if (ctx.get.column_ids[i] == id) {
			return i;
		}
	}
	ctx.get.column_ids.push_back(id);
	return ctx.get.column_ids.size() - 1;
}

static void RewriteExpression(unique_ptr<Expression> &expr, RewriteCtx &ctx, vector<idx_t> &added) {
	ComputedColumn c;
	if (MatchComputed(*expr, ctx, c)) {
		const idx_t before = ctx.get.column_ids.size();
		const idx_t pos = ComputedBinding(ctx, c);
		if (pos != DConstants::INVALID_INDEX) {
			if (ctx.get.column_ids.size() != before) {
				added.push_back(pos);
			}
			auto ref = make_uniq<BoundColumnRefExpression>(c.name, c.type, ColumnBinding(ctx.get.table_index, pos));
			ref->alias = expr->alias;
			expr = std::move(ref);
			ctx.rewritten++;
			return;
		}
	}
	ExpressionIterator::EnumerateChildren(*expr, [&](unique_ptr<Expression> &child) { RewriteExpression(child, ctx, added); });
}

static void CountReferences(const Expression &e, idx_t table_index, vector<idx_t> &refs) {
	if (e.expression_class == ExpressionClass::BOUND_COLUMN_REF) {
		auto &ref = e.Cast<BoundColumnRefExpression>();
		if (ref.binding.table_index == table_index && ref.binding.column_index < refs.size()) {
			refs[ref.binding.column_index]++;
		}
	}
	ExpressionIterator::EnumerateChildren(e, [&](const Expression &child) { CountReferences(child, table_index, refs); });
}

static bool IsExonScan(LogicalOperator &op) {
	if (op.type != LogicalOperatorType::LOGICAL_GET) {
		return false;
	}
	auto &get = op.Cast<LogicalGet>();
	return get.function.function == ScanFunction && get.bind_data && get.projection_ids.empty() && get.children.empty();
}

static void OptimizeOperator(LogicalOperator &op) {
	for (auto &child : op.children) {
		OptimizeOperator(*child);
	}
	// `op` consumes a scan directly, or through a chain of filters that pass the scan's columns up unchanged
	for (auto &child : op.children) {
		vector<LogicalFilter *> filters;
		LogicalOperator *cur = child.get();
		while (cur->type == LogicalOperatorType::LOGICAL_FILTER && cur->children.size() == 1) {
			filters.push_back(&cur->Cast<LogicalFilter>());
			cur = cur->children[0].get();
		}
		if (!IsExonScan(*cur) || op.type == LogicalOperatorType::LOGICAL_FILTER) {
			continue; // a filter is handled from the operator above it
		}
		auto &get = cur->Cast<LogicalGet>();
		auto &bind = get.bind_data->Cast<ScanBindData>();
		RewriteCtx ctx {get, bind, filters.empty()};
		vector<idx_t> added; // positions in get.column_ids created by this rewrite
		LogicalOperatorVisitor::EnumerateExpressions(op, [&](unique_ptr<Expression> *e) { RewriteExpression(*e, ctx, added); });
		for (auto f : filters) {
			RewriteCtx fctx {get, bind, false};
			LogicalOperatorVisitor::EnumerateExpressions(*f, [&](unique_ptr<Expression> *e) { RewriteExpression(*e, fctx, added); });
			ctx.rewritten += fctx.rewritten;
		}
		if (ctx.rewritten == 0) {
			continue;
		}
		// filters that pass only some of their child's columns up must pass the new ones too (bottom filter first)
		for (idx_t fi = filters.size(); fi-- > 0;) {
			auto &f = *filters[fi];
			if (f.projection_map.empty()) {
				continue; // passes everything
			}
			auto below = f.children[0]->GetColumnBindings();
			for (idx_t pos : added) {
				for (idx_t b = 0; b < below.size(); b++) {
					if (below[b].table_index == get.table_index && below[b].column_index == pos) {
						f.projection_map.push_back(b);
					}
				}
			}
		}
		// file columns that nothing reads any more stay in column_ids (table filters and bindings index it) but are not
		// materialised: the scan hands DuckDB a constant NULL vector for them.  Only when `op` ends the columns' life.
		if (op.type == LogicalOperatorType::LOGICAL_PROJECTION || op.type == LogicalOperatorType::LOGICAL_AGGREGATE_AND_GROUP_BY) {
			vector<idx_t> refs(get.column_ids.size(), 0);
			LogicalOperatorVisitor::EnumerateExpressions(op, [&](unique_ptr<Expression> *e) { CountReferences(**e, get.table_index, refs); });
			for (auto f : filters) {
				LogicalOperatorVisitor::EnumerateExpressions(*f, [&](unique_ptr<Expression> *e) { CountReferences(**e, get.table_index, refs); });
			}
			for (idx_t i = 0; i < get.column_ids.size(); i++) {
				const auto id = get.column_ids[i];
				if (id != COLUMN_IDENTIFIER_ROW_ID && id < bind.names.size() && refs[i] == 0) {
					bind.dead[id] = true;
				}
			}
		}
	}
}

static void ExonOptimize(ClientContext &context, OptimizerExtensionInfo *info, unique_ptr<LogicalOperator> &plan) {
	OptimizeOperator(*plan);
}

static void RegisterScan(ClientContext &context, const string &name, const string &file_type) {
	TableFunction scan(name, {LogicalType::VARCHAR}, ScanFunction, ScanBind, ScanInitGlobal, ScanInitLocal);
	scan.cardinality = ScanCardinality;        // module.cpp:307
	scan.get_batch_index = ScanGetBatchIndex;  // module.cpp:308
	scan.table_scan_progress = ScanProgress;
	scan.function_info = make_shared<ScanInfo>(file_type);
	scan.named_parameters["compression"] = LogicalType::VARCHAR;
	scan.named_parameters["gpus"] = LogicalType::INTEGER; // extension of this build: device pipelines (0 = automatic)
	scan.projection_pushdown = true;
	scan.filter_pushdown = true;
	scan.pushdown_complex_filter = ScanPushdownComplexFilter;
	scan.to_string = ScanToString;
	auto &catalog = Catalog::GetSystemCatalog(context);
	CreateTableFunctionInfo info(scan);
	catalog.CreateTableFunction(context, &info);
}

// ---- exon_gpu_stats(): per-scan counters of the readers that ran in this process (SURVEY 5 "Metrics / logging": the reference has
// only the lines_read atomic of module.cpp:40,276).  One row per closed reader, newest first: what was read, from where and how
// (I/O path, compression), rows, and seconds per pipeline stage.
struct StatsState : public GlobalTableFunctionState {
	bool done = false;
};
static unique_ptr<FunctionData> StatsBind(ClientContext &, TableFunctionBindInput &, vector<LogicalType> &types, vector<string> &names) {
	const vector<std::pair<string, LogicalType>> cols = {
	    {"path", LogicalType::VARCHAR},          {"files", LogicalType::INTEGER},        {"format", LogicalType::VARCHAR},
	    {"device", LogicalType::INTEGER},        {"compression", LogicalType::VARCHAR},  {"io_path", LogicalType::VARCHAR},
	    {"failed", LogicalType::BOOLEAN},        {"file_bytes", LogicalType::BIGINT},    {"bytes_done", LogicalType::BIGINT},
	    {"rows", LogicalType::BIGINT},           {"blocks", LogicalType::BIGINT},        {"seconds_total", LogicalType::DOUBLE},
	    {"seconds_io", LogicalType::DOUBLE},     {"seconds_device", LogicalType::DOUBLE}, {"seconds_scan", LogicalType::DOUBLE},
	    {"seconds_select", LogicalType::DOUBLE}, {"seconds_materialise", LogicalType::DOUBLE}, {"seconds_first_block", LogicalType::DOUBLE},
	    {"gb_per_s", LogicalType::DOUBLE}};
	for (auto &c : cols) {
		names.push_back(c.first);
		types.push_back(c.second);
	}
	return nullptr;
}
static unique_ptr<GlobalTableFunctionState> StatsInit(ClientContext &, TableFunctionInitInput &) {
	return make_uniq<StatsState>();
}
static void StatsFunction(ClientContext &, TableFunctionInput &input, DataChunk &output) {
	auto &state = input.global_state->Cast<StatsState>();
	if (state.done) {
		output.SetCardinality(0);
		return;
	}
	state.done = true;
	exb_scan_stats st[64];
	int n = 0;
	exb_stats_snapshot(st, 64, &n);
	static const char *comp[] = {"none", "gzip (zlib stream)", "zstd", "bzip2", "xz", "bgzf (inflated on the device)"};
	for (int i = 0; i < n; i++) {
		const exb_scan_stats &r = st[i];
		idx_t c = 0;
		output.SetValue(c++, i, Value(string(r.path)));
		output.SetValue(c++, i, Value::INTEGER(r.n_files));
		output.SetValue(c++, i, Value(r.format == 1 ? "fasta" : "fastq"));
		output.SetValue(c++, i, Value::INTEGER(r.device));
		output.SetValue(c++, i, Value(comp[r.compression >= 0 && r.compression <= 5 ? r.compression : 0]));
		output.SetValue(c++, i, Value(r.io_path ? "registered page cache" : "pinned blocks"));
		output.SetValue(c++, i, Value::BOOLEAN(r.failed != 0));
		output.SetValue(c++, i, Value::BIGINT(r.file_bytes));
		output.SetValue(c++, i, Value::BIGINT(r.bytes_done));
		output.SetValue(c++, i, Value::BIGINT(r.rows));
		output.SetValue(c++, i, Value::BIGINT(r.blocks));
		output.SetValue(c++, i, Value::DOUBLE(r.seconds_total));
		output.SetValue(c++, i, Value::DOUBLE(r.seconds_io));
		output.SetValue(c++, i, Value::DOUBLE(r.seconds_device));
		output.SetValue(c++, i, Value::DOUBLE(r.seconds_scan));
		output.SetValue(c++, i, Value::DOUBLE(r.seconds_select));
		output.SetValue(c++, i, Value::DOUBLE(r.seconds_materialise));
		output.SetValue(c++, i, Value::DOUBLE(r.seconds_first_block));
		output.SetValue(c++, i, Value::DOUBLE(r.seconds_total > 0 ? (double)r.bytes_done / r.seconds_total / 1e9 : 0.0));
	}
	output.SetCardinality(n);
}
static void RegisterStats(ClientContext &context) {
	TableFunction fn("exon_gpu_stats", {}, StatsFunction, StatsBind, StatsInit);
	auto &catalog = Catalog::GetSystemCatalog(context);
	CreateTableFunctionInfo info(fn);
	catalog.CreateTableFunction(context, &info);
}

static unique_ptr<TableRef> ExonReplacementScan(ClientContext &context, const string &table_name, ReplacementScanData *data) {
	auto lower_name = StringUtil::Lower(table_name);
	auto res = replacement_scan(lower_name.c_str());
	if (!res.file_type) {
		return nullptr;
	}
	const string file_type(res.file_type);
	exb_free_string(res.file_type);
	auto table_function = make_uniq<TableFunctionRef>();
	vector<unique_ptr<ParsedExpression>> children;
	children.push_back(make_uniq<ConstantExpression>(Value(table_name)));
	if (file_type == "FASTA") {
		table_function->function = make_uniq<FunctionExpression>("read_fasta", std::move(children));
	} else if (file_type == "FASTQ") {
		table_function->function = make_uniq<FunctionExpression>("read_fastq", std::move(children));
	} else {
		return nullptr;
	}
	return std::move(table_function);
}

// ------------------------------------------------------------------ scalar functions
// Rows of one DataChunk packed as a contiguous host column (NULL rows are skipped: `rows` maps packed -> chunk row).
struct Packed {
	vector<int64_t> offsets;
	vector<uint8_t> data;
	vector<idx_t> rows;
};

static void PackStrings(Vector &input, idx_t count, Packed &p, UnifiedVectorFormat &fmt) {
	input.ToUnifiedFormat(count, fmt);
	auto strings = UnifiedVectorFormat::GetData<string_t>(fmt);
	p.offsets.clear();
	p.rows.clear();
	p.offsets.push_back(0);
	int64_t total = 0;
	for (idx_t i = 0; i < count; i++) {
		auto idx = fmt.sel->get_index(i);
		if (fmt.validity.RowIsValid(idx)) {
			total += strings[idx].GetSize();
		}
	}
	p.data.resize((size_t)total + 16);
	int64_t at = 0;
	for (idx_t i = 0; i < count; i++) {
		auto idx = fmt.sel->get_index(i);
		if (!fmt.validity.RowIsValid(idx)) {
			continue;
		}
		const auto &s = strings[idx];
		memcpy(p.data.data() + at, s.GetData(), s.GetSize());
		at += s.GetSize();
		p.offsets.push_back(at);
		p.rows.push_back(i);
	}
}

static void FinishResult(Vector &result, DataChunk &args, idx_t count) {
	if (args.AllConstant()) {
		result.SetVectorType(VectorType::CONSTANT_VECTOR);
	}
}

// gc_content(VARCHAR) -> FLOAT: per row (float)#{G,C} / (float)len, '' -> 0, NULL -> NULL
// (sequence_functions/module.cpp:131-158; the reference's chunk-collapse bug -- SURVEY finding 4 -- is not reproduced)
static void GcContentFunction(DataChunk &args, ExpressionState &state, Vector &result) {
	const idx_t count = args.size();
	Packed p;
	UnifiedVectorFormat fmt;
	PackStrings(args.data[0], count, p, fmt);
	vector<float> out(p.rows.size() + 1);
	if (exb_gc_content_host(p.offsets.data(), p.data.data(), (int64_t)p.rows.size(), out.data()) != 0) {
		throw InvalidInputException(exb_last_error());
	}
	result.SetVectorType(VectorType::FLAT_VECTOR);
	auto res = FlatVector::GetData<float>(result);
	auto &validity = FlatVector::Validity(result);
	for (idx_t i = 0; i < count; i++) {
		validity.SetInvalid(i);
	}
	for (idx_t k = 0; k < p.rows.size(); k++) {
		res[p.rows[k]] = out[k];
		validity.SetValid(p.rows[k]);
	}
	FinishResult(result, args, count);
}

// reverse_complement / complement: byte LUT with the reference's tables (module.cpp:30-69 / 81-121);
// a byte outside ACGT raises InvalidInputException("Invalid character in sequence: <c>")
template <int MODE>
static void SeqMapFunction(DataChunk &args, ExpressionState &state, Vector &result) {
	const idx_t count = args.size();
	Packed p;
	UnifiedVectorFormat fmt;
	PackStrings(args.data[0], count, p, fmt);
	const int64_t nb = p.offsets.back();
	vector<uint8_t> out((size_t)nb + 16);
	int64_t bad = -1;
	if (exb_seq_map_host(p.data.data(), nb, MODE, out.data(), &bad) != 0) {
		throw InvalidInputException(exb_last_error());
	}
	if (bad >= 0) {
		throw InvalidInputException("Invalid character in sequence: " + string(1, (char)p.data[(size_t)bad]));
	}
	result.SetVectorType(VectorType::FLAT_VECTOR);
	auto res = FlatVector::GetData<string_t>(result);
	auto &validity = FlatVector::Validity(result);
	for (idx_t i = 0; i < count; i++) {
		validity.SetInvalid(i);
	}
	for (idx_t k = 0; k < p.rows.size(); k++) {
		const int64_t b = p.offsets[k], e = p.offsets[k + 1];
		res[p.rows[k]] = StringVector::AddString(result, reinterpret_cast<const char *>(out.data()) + b, (idx_t)(e - b));
		validity.SetValid(p.rows[k]);
	}
	FinishResult(result, args, count);
}

// translate_dna_to_aa(VARCHAR) -> VARCHAR (module.cpp:260-360): standard codon table; "Invalid sequence length: <n>" /
// "Invalid codon: <xyz>" for the first offending row, the length checked before the codons as the reference does
static void TranslateFunction(DataChunk &args, ExpressionState &state, Vector &result) {
	const idx_t count = args.size();
	Packed p;
	UnifiedVectorFormat fmt;
	PackStrings(args.data[0], count, p, fmt);
	const int64_t nb = p.offsets.back();
	vector<uint8_t> out((size_t)nb / 3 + 16);
	int64_t status[2] = {-1, -1};
	if (exb_translate_host(p.offsets.data(), p.data.data(), (int64_t)p.rows.size(), out.data(), status) != 0) {
		throw InvalidInputException(exb_last_error());
	}
	if (status[1] >= 0) {
		throw InvalidInputException("Invalid codon: " + string(reinterpret_cast<const char *>(p.data.data()) + status[1], 3));
	}
	if (status[0] >= 0) {
		throw InvalidInputException("Invalid sequence length: " + std::to_string(p.offsets[status[0] + 1] - p.offsets[status[0]]));
	}
	result.SetVectorType(VectorType::FLAT_VECTOR);
	auto res = FlatVector::GetData<string_t>(result);
	auto &validity = FlatVector::Validity(result);
	for (idx_t i = 0; i < count; i++) {
		validity.SetInvalid(i);
	}
	for (idx_t k = 0; k < p.rows.size(); k++) {
		const int64_t b = p.offsets[k] / 3, e = p.offsets[k + 1] / 3;
		res[p.rows[k]] = StringVector::AddString(result, reinterpret_cast<const char *>(out.data()) + b, (idx_t)(e - b));
		validity.SetValid(p.rows[k]);
	}
	FinishResult(result, args, count);
}

// quality_score_string_to_list(VARCHAR) -> INTEGER[]: (signed char)c - 33 per byte (fastq_functions/module.cpp:32-50).
// Deviation, documented in DESIGN.md: '' gives [] and NULL gives NULL, where the reference raises INTERNAL errors.
static void QualityToListFunction(DataChunk &args, ExpressionState &state, Vector &result) {
	const idx_t count = args.size();
	Packed p;
	UnifiedVectorFormat fmt;
	PackStrings(args.data[0], count, p, fmt);
	const int64_t nb = p.offsets.back();
	result.SetVectorType(VectorType::FLAT_VECTOR);
	ListVector::Reserve(result, (idx_t)nb);
	auto &child = ListVector::GetEntry(result);
	auto child_data = FlatVector::GetData<int32_t>(child);
	if (exb_quality_decode_host(p.data.data(), nb, child_data) != 0) {
		throw InvalidInputException(exb_last_error());
	}
	ListVector::SetListSize(result, (idx_t)nb);
	auto entries = FlatVector::GetData<list_entry_t>(result);
	auto &validity = FlatVector::Validity(result);
	for (idx_t i = 0; i < count; i++) {
		validity.SetInvalid(i);
		entries[i].offset = 0;
		entries[i].length = 0;
	}
	for (idx_t k = 0; k < p.rows.size(); k++) {
		entries[p.rows[k]].offset = (uint64_t)p.offsets[k];
		entries[p.rows[k]].length = (uint64_t)(p.offsets[k + 1] - p.offsets[k]);
		validity.SetValid(p.rows[k]);
	}
	FinishResult(result, args, count);
}

// ------------------------------------------------------------------ COPY ... TO (FORMAT 'fastq' | 'fasta')
// The writers the reference declares (fastq_functions/module.hpp:30 GetFastqCopyFunction) and removed; their statements
// survive, commented out, in test_fastq_copy.test / test_fasta_copy.test and are what this mirrors:
//   COPY (query) TO 'file' (FORMAT 'fastq' | 'fasta' [, COMPRESSION 'gzip' | 'zstd'] [, FORCE true])
// compression defaults to the file suffix (.gz / .zst); an existing file is an error unless FORCE is given.  The query
// supplies the reader's columns by position: (name | id, description, sequence[, quality_scores]), all VARCHAR.
// Rows are staged by exb_writer_append and formatted on the GPU (csrc/writer_ops.cu); there is no CPU formatter.
struct CopyBindData : public FunctionData {
	string file_type;
	string compression; // "" = by suffix of the ORIGINAL path (DuckDB may hand init_global a tmp_ name)
	bool force = false;
	int32_t line_width = 80;
	idx_t n_cols = 0;

	unique_ptr<FunctionData> Copy() const override {
		return make_uniq<CopyBindData>(*this);
	}
	bool Equals(const FunctionData &other_p) const override {
		auto &o = other_p.Cast<CopyBindData>();
		return file_type == o.file_type && compression == o.compression && force == o.force && line_width == o.line_width;
	}
};
struct CopyGlobalState : public GlobalFunctionData {
	std::mutex lock;
	exb_writer *writer = nullptr;
	~CopyGlobalState() override {
		if (writer) {
			exb_writer_close(writer, nullptr, nullptr);
		}
	}
};
struct CopyLocalState : public LocalFunctionData {};

template <bool FASTA>
static unique_ptr<FunctionData> CopyBind(ClientContext &context, CopyInfo &info, vector<string> &names, vector<LogicalType> &sql_types) {
	auto bind = make_uniq<CopyBindData>();
	bind->file_type = FASTA ? "fasta" : "fastq";
	bind->n_cols = FASTA ? 3 : 4;
	for (auto &option : info.options) {
		auto key = StringUtil::Lower(option.first);
		if (key == "compression") {
			if (option.second.size() != 1) {
				throw BinderException("COMPRESSION needs one value: 'gzip', 'zstd' or 'none'");
			}
			bind->compression = StringUtil::Lower(option.second[0].ToString());
		} else if (key == "force") {
			bind->force = option.second.empty() || option.second[0].CastAs(context, LogicalType::BOOLEAN).GetValue<bool>();
		} else if (key == "line_width" && FASTA) {
			if (option.second.size() != 1) {
				throw BinderException("LINE_WIDTH needs one value");
			}
			bind->line_width = option.second[0].CastAs(context, LogicalType::INTEGER).GetValue<int32_t>();
			if (bind->line_width < 1) {
				throw BinderException("LINE_WIDTH must be positive");
			}
		} else {
			throw BinderException("Unknown option for COPY ... TO (FORMAT '%s'): %s", bind->file_type, option.first);
		}
	}
	if (sql_types.size() != bind->n_cols) {
		throw BinderException("COPY ... TO (FORMAT '%s') needs %llu columns (%s), the query has %llu", bind->file_type, bind->n_cols,
		                      FASTA ? "id, description, sequence" : "name, description, sequence, quality_scores", sql_types.size());
	}
	for (idx_t i = 0; i < sql_types.size(); i++) {
		if (sql_types[i].id() != LogicalTypeId::VARCHAR) {
			throw BinderException("COPY ... TO (FORMAT '%s'): column %s must be VARCHAR", bind->file_type, names[i]);
		}
	}
	if (bind->compression.empty()) { // the suffix of the path the user wrote
		auto lower = StringUtil::Lower(info.file_path);
		bind->compression = StringUtil::EndsWith(lower, ".gz") ? "gzip" : (StringUtil::EndsWith(lower, ".zst") ? "zstd" : "none");
	}
	// the reference's writer refused to replace a file (test_fasta_copy.test:43-50); DuckDB itself would write a tmp_ file
	// and move it over the old one, so the check has to happen here, on the path the user wrote
	if (!bind->force && info.file_path != "/dev/stdout" && FileSystem::GetFileSystem(context).FileExists(info.file_path)) {
		throw IOException("%s exists; COPY ... TO (FORMAT '%s', FORCE true) replaces it", info.file_path, bind->file_type);
	}
	return std::move(bind);
}

static unique_ptr<GlobalFunctionData> CopyInitGlobal(ClientContext &context, FunctionData &bind_data, const string &file_path) {
	auto &bind = bind_data.Cast<CopyBindData>();
	auto state = make_uniq<CopyGlobalState>();
	// the bind-time check already settled whether the target may be replaced; a tmp_ file left by a failed run may too
	if (exb_writer_open(file_path.c_str(), bind.file_type.c_str(), bind.compression.c_str(), 1, 0, &state->writer) != 0) {
		throw IOException(exb_last_error());
	}
	if (bind.file_type == "fasta" && exb_writer_set_line_width(state->writer, bind.line_width) != 0) {
		throw IOException(exb_last_error());
	}
	return std::move(state);
}

static unique_ptr<LocalFunctionData> CopyInitLocal(ExecutionContext &context, FunctionData &bind_data) {
	return make_uniq<CopyLocalState>();
}

static void CopySink(ExecutionContext &context, FunctionData &bind_data, GlobalFunctionData &gstate, LocalFunctionData &lstate,
                     DataChunk &input) {
	auto &bind = bind_data.Cast<CopyBindData>();
	auto &state = gstate.Cast<CopyGlobalState>();
	const idx_t count = input.size();
	if (count == 0) {
		return;
	}
	// every column as contiguous offsets + bytes; NULL counts as empty (and, for the description, as "no description")
	vector<int64_t> offsets[4];
	vector<uint8_t> data[4];
	vector<uint8_t> desc_valid(count, 1);
	for (idx_t c = 0; c < bind.n_cols; c++) {
		UnifiedVectorFormat fmt;
		input.data[c].ToUnifiedFormat(count, fmt);
		auto strings = UnifiedVectorFormat::GetData<string_t>(fmt);
		offsets[c].resize(count + 1);
		int64_t total = 0;
		for (idx_t i = 0; i < count; i++) {
			auto idx = fmt.sel->get_index(i);
			offsets[c][i] = total;
			if (fmt.validity.RowIsValid(idx)) {
				total += strings[idx].GetSize();
			} else if (c == 1) {
				desc_valid[i] = 0;
			} else if (c == 0) {
				throw InvalidInputException("COPY ... TO (FORMAT '%s'): %s is NULL", bind.file_type, c == 0 && bind.n_cols == 3 ? "id" : "name");
			}
		}
		offsets[c][count] = total;
		data[c].resize((size_t)total + 16);
		for (idx_t i = 0; i < count; i++) {
			auto idx = fmt.sel->get_index(i);
			if (fmt.validity.RowIsValid(idx)) {
				memcpy(data[c].data() + offsets[c][i], strings[idx].GetData(), strings[idx].GetSize());
			}
		}
	}
	const int64_t *off_p[4] = {offsets[0].data(), offsets[1].data(), offsets[2].data(), bind.n_cols == 4 ? offsets[3].data() : nullptr};
	const uint8_t *data_p[4] = {data[0].data(), data[1].data(), data[2].data(), bind.n_cols == 4 ? data[3].data() : nullptr};
	std::lock_guard<std::mutex> guard(state.lock);
	if (exb_writer_append(state.writer, (int64_t)count, off_p, data_p, desc_valid.data()) != 0) {
		throw IOException(exb_last_error());
	}
}

static void CopyCombine(ExecutionContext &context, FunctionData &bind_data, GlobalFunctionData &gstate, LocalFunctionData &lstate) {
}

static void CopyFinalize(ClientContext &context, FunctionData &bind_data, GlobalFunctionData &gstate) {
	auto &state = gstate.Cast<CopyGlobalState>();
	std::lock_guard<std::mutex> guard(state.lock);
	auto writer = state.writer;
	state.writer = nullptr;
	if (writer && exb_writer_close(writer, nullptr, nullptr) != 0) {
		throw IOException(exb_last_error());
	}
}

// one writer, rows in the order the sink sees them: never PARALLEL / BATCH copy (the default when this is null would
// depend on preserve_insertion_order)
static CopyFunctionExecutionMode CopyExecutionMode(bool preserve_insertion_order, bool supports_batch_index) {
	return CopyFunctionExecutionMode::REGULAR_COPY_TO_FILE;
}

template <bool FASTA>
static void RegisterCopy(ClientContext &context) {
	CopyFunction fn(FASTA ? "fasta" : "fastq");
	fn.copy_to_bind = CopyBind<FASTA>;
	fn.copy_to_initialize_global = CopyInitGlobal;
	fn.copy_to_initialize_local = CopyInitLocal;
	fn.copy_to_sink = CopySink;
	fn.copy_to_combine = CopyCombine;
	fn.copy_to_finalize = CopyFinalize;
	fn.execution_mode = CopyExecutionMode;
	fn.extension = FASTA ? "fasta" : "fastq";
	CreateCopyFunctionInfo info(fn);
	Catalog::GetSystemCatalog(context).CreateCopyFunction(context, info);
}

static void RegisterScalar(ClientContext &context, const string &name, const LogicalType &ret, scalar_function_t fn) {
	ScalarFunctionSet set(name);
	set.AddFunction(ScalarFunction({LogicalType::VARCHAR}, ret, std::move(fn)));
	CreateScalarFunctionInfo info(set);
	Catalog::GetSystemCatalog(context).CreateFunction(context, info);
}

static void LoadInternal(DatabaseInstance &instance) {
	Connection con(instance);
	con.BeginTransaction();
	auto &context = *con.context;
	auto &config = DBConfig::GetConfig(context);
	RegisterScalar(context, "gc_content", LogicalType::FLOAT, GcContentFunction);
	RegisterScalar(context, "reverse_complement", LogicalType::VARCHAR, SeqMapFunction<EXB_MAP_REVERSE_COMPLEMENT>);
	RegisterScalar(context, "complement", LogicalType::VARCHAR, SeqMapFunction<EXB_MAP_COMPLEMENT>);
	RegisterScalar(context, "transcribe", LogicalType::VARCHAR, SeqMapFunction<EXB_MAP_TRANSCRIBE>);
	RegisterScalar(context, "reverse_transcribe", LogicalType::VARCHAR, SeqMapFunction<EXB_MAP_REVERSE_TRANSCRIBE>);
	RegisterScalar(context, "translate_dna_to_aa", LogicalType::VARCHAR, TranslateFunction);
	RegisterScalar(context, "quality_score_string_to_list", LogicalType::LIST(LogicalType::INTEGER), QualityToListFunction);
	RegisterScan(context, "read_fasta", "fasta");
	RegisterScan(context, "read_fastq", "fastq");
	RegisterCopy<false>(context);
	RegisterCopy<true>(context);
	RegisterStats(context);
	config.replacement_scans.emplace_back(ExonReplacementScan);
	OptimizerExtension fuse;
	fuse.optimize_function = ExonOptimize;
	config.optimizer_extensions.push_back(fuse);
	con.Commit();
}

} // namespace exon_b200

extern "C" {
DUCKDB_EXTENSION_API void exon_init(duckdb::DatabaseInstance &db) {
	exon_b200::LoadInternal(db);
}
DUCKDB_EXTENSION_API const char *exon_version() {
	return duckdb::DuckDB::LibraryVersion();
}
}
