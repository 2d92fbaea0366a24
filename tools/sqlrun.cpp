// sqlrun -- minimal DuckDB host for the extension tests.
//
// Links the libduckdb the extension was compiled against (DuckDB's extension
// ABI is C++, and exon_version() must equal DuckDB::LibraryVersion(),
// duckdb/src/main/extension/extension_load.cpp:163-258), enables unsigned
// extensions like the reference's test runner does
// (duckdb/test/helpers/test_helpers.cpp:139), reads ';'-terminated statements
// from stdin and prints ONE JSON line per statement:
//   {"ok": true, "names": [...], "types": [...], "rows": [[...], ...]}
//   {"ok": false, "error": "..."}
// Values are rendered with Value::ToString(); NULL becomes JSON null.
//
// usage: sqlrun [-threads N] < script.sql
#include <iostream>
#include <chrono>
#include <sstream>
#include <thread>
#include <string>

#include "duckdb.hpp"

using namespace duckdb;

static std::string json_escape(const std::string &s) {
	std::string o;
	o.reserve(s.size() + 2);
	o.push_back('"');
	for (unsigned char c : s) {
		switch (c) {
		case '"': o += "\\\""; break;
		case '\\': o += "\\\\"; break;
		case '\n': o += "\\n"; break;
		case '\r': o += "\\r"; break;
		case '\t': o += "\\t"; break;
		default:
			if (c < 0x20) {
				char buf[8];
				snprintf(buf, sizeof(buf), "\\u%04x", c);
				o += buf;
			} else {
				o.push_back((char)c);
			}
		}
	}
	o.push_back('"');
	return o;
}

int main(int argc, char **argv) {
	int threads = 0;
	for (int i = 1; i + 1 < argc; i++)
		if (std::string(argv[i]) == "-threads") threads = atoi(argv[i + 1]);
	DBConfig config;
	config.options.allow_unsigned_extensions = true;
	if (threads > 0) config.options.maximum_threads = threads;
	DuckDB db(nullptr, &config);
	Connection con(db);

	std::stringstream ss;
	ss << std::cin.rdbuf();
	const std::string text = ss.str();
	// split on ';' outside single quotes; '--' / '#' comment lines are dropped
	std::string stmt;
	bool in_str = false;
	size_t i = 0;
	auto run = [&](const std::string &sql) {
		bool blank = true;
		for (char c : sql)
			if (!isspace((unsigned char)c)) blank = false;
		if (blank) return;
		{  // ".sleep <ms>": a pause between two statements (an interactive user's think time); prints nothing
			size_t a = 0;
			while (a < sql.size() && isspace((unsigned char)sql[a])) a++;
			if (sql.compare(a, 6, ".sleep") == 0) {
				std::this_thread::sleep_for(std::chrono::milliseconds(atoi(sql.c_str() + a + 6)));
				return;
			}
		}
		const auto t0 = std::chrono::steady_clock::now();
		auto res = con.Query(sql);
		if (res->HasError()) {
			std::cout << "{\"ok\": false, \"error\": " << json_escape(res->GetError()) << "}" << std::endl;
			return;
		}
		std::cout << "{\"ok\": true, \"names\": [";
		for (idx_t c = 0; c < res->ColumnCount(); c++) std::cout << (c ? ", " : "") << json_escape(res->names[c]);
		std::cout << "], \"types\": [";
		for (idx_t c = 0; c < res->ColumnCount(); c++) std::cout << (c ? ", " : "") << json_escape(res->types[c].ToString());
		std::cout << "], \"rows\": [";
		bool first = true;
		while (true) {
			auto chunk = res->Fetch();
			if (!chunk || chunk->size() == 0) break;
			for (idx_t r = 0; r < chunk->size(); r++) {
				std::cout << (first ? "[" : ", [");
				first = false;
				for (idx_t c = 0; c < chunk->ColumnCount(); c++) {
					Value v = chunk->GetValue(c, r);
					std::cout << (c ? ", " : "") << (v.IsNull() ? std::string("null") : json_escape(v.ToString()));
				}
				std::cout << "]";
			}
		}
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		std::cout << "], \"ms\": " << ms << "}" << std::endl;
	};
	while (i < text.size()) {
		if (!in_str && (stmt.empty() || stmt.back() == '\n')) {
			// comment line?
			size_t j = i;
			while (j < text.size() && (text[j] == ' ' || text[j] == '\t')) j++;
			if (j < text.size() && (text[j] == '#' || (text[j] == '-' && j + 1 < text.size() && text[j + 1] == '-'))) {
				while (i < text.size() && text[i] != '\n') i++;
				continue;
			}
		}
		const char c = text[i++];
		if (c == '\'') in_str = !in_str;
		if (c == ';' && !in_str) {
			run(stmt);
			stmt.clear();
			continue;
		}
		stmt.push_back(c);
	}
	run(stmt);
	return 0;
}
