// ref_loader.cpp -- TEST INFRASTRUCTURE.  Loader for the reference's OWN glue.
//
// oracle/build_ref.sh compiles the reference's unmodified sources
//   exon/src/exon/arrow_table_function/module.cpp   (read_fasta / read_fastq table function, replacement scan)
//   exon/src/exon/sequence_functions/module.cpp     (gc_content, reverse_complement, complement, ...)
//   exon/src/exon/fastq_functions/module.cpp        (quality_score_string_to_list)
// from where they lie under /root/reference, together with this file, into
// oracle/_ref/exon.duckdb_extension.  The reference's own entry point
// (exon/src/exon_extension.cpp:25-96) also registers SAM/BAM/VCF/GFF/... functions
// whose Rust symbols are outside the hot path, so this loader registers exactly
// the path's functions the same way that file does (same calls, lines 40-55,81).
//
// The Rust staticlib's `new_reader` / `replacement_scan` (exon/include/rust.hpp:41-48)
// are resolved from libexon_b200.so -- the drop-in boundary: the reference's C++
// glue runs unchanged on top of the B200 engine.  The scalar functions in this
// build are the REFERENCE's CPU implementations: the live oracle for SURVEY 8a rows a9-a12.
#define DUCKDB_EXTENSION_MAIN
#include "duckdb.hpp"
#include "duckdb/main/extension_util.hpp"
#include "exon/arrow_table_function/module.hpp"
#include "exon/fastq_functions/module.hpp"
#include "exon/sequence_functions/module.hpp"

using namespace duckdb;

static void LoadPath(DatabaseInstance &instance) {
	Connection con(instance);
	con.BeginTransaction();
	auto &context = *con.context;
	auto &catalog = Catalog::GetSystemCatalog(context);
	auto &config = DBConfig::GetConfig(context);
	auto sequence_functions = exon::SequenceFunctions::GetSequenceFunctions();
	for (auto &fun : sequence_functions) {
		catalog.CreateFunction(context, fun);
	}
	exon::WTArrowTableFunction::Register("read_fasta", "fasta", context);
	exon::WTArrowTableFunction::Register("read_fastq", "fastq", context);
	auto q = exon::FastqFunctions::GetQualityScoreStringToList();
	catalog.CreateFunction(context, *q);
	config.replacement_scans.emplace_back(exon::WTArrowTableFunction::ReplacementScan);
	con.Commit();
}

extern "C" {
DUCKDB_EXTENSION_API void exon_init(duckdb::DatabaseInstance &db) {
	LoadPath(db);
}
DUCKDB_EXTENSION_API const char *exon_version() {
	return duckdb::DuckDB::LibraryVersion();
}
}
