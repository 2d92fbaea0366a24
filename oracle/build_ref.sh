#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.  Builds, in THIS container only (needs /root/reference):
#   build/rt/libduckdb.so, build/rt/sqlrun     the vendored DuckDB v0.8.1 (duckdb submodule of the reference) as a
#                                              shared library + our minimal SQL host (tools/sqlrun.cpp)
#   oracle/_ref/exon.duckdb_extension          the reference's unmodified C++ glue + scalar functions
#                                              (sources compiled where they lie), see oracle/ref_loader.cpp
# Nothing is copied out of /root/reference; outputs are git-ignored and travel to the GPU box with the snapshot.
# The vendored DuckDB is a CMake project: it is configured out of tree (the reference's own top-level CMakeLists
# cannot configure offline -- FetchContent of arrow/httplib/json/Corrosion, SURVEY 0.2 -- and is NOT used).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=/root/reference
[ -d $REF/duckdb/src/include ] || { echo "no /root/reference: nothing to build"; exit 0; }
mkdir -p $ROOT/build/rt $ROOT/oracle/_ref
if [ ! -f $ROOT/build/rt/libduckdb.so ]; then
    cmake -G Ninja -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_BUILD_TYPE=Release -DBUILD_UNITTESTS=OFF -DBUILD_SHELL=ON \
        -DENABLE_SANITIZER=OFF -DENABLE_UBSAN=OFF -DBUILD_PARQUET_EXTENSION=OFF -DBUILD_JEMALLOC_EXTENSION=OFF \
        -S $REF/duckdb -B $ROOT/build/duckdb > $ROOT/build/cmake.log 2>&1
    ninja -C $ROOT/build/duckdb -j8 src/libduckdb.so > $ROOT/build/ninja.log 2>&1
    cp $ROOT/build/duckdb/src/libduckdb.so $ROOT/build/rt/libduckdb.so
    strip $ROOT/build/rt/libduckdb.so
    rm -rf $ROOT/build/duckdb
fi
CXXF="-O2 -std=c++17 -fPIC -w -I$REF/duckdb/src/include -I$REF/duckdb/third_party/re2 -I$REF/duckdb/third_party/fmt/include -I$REF/duckdb/third_party/utf8proc/include"
if [ ! -f $ROOT/build/rt/sqlrun ] || [ $ROOT/tools/sqlrun.cpp -nt $ROOT/build/rt/sqlrun ]; then
    g++ $CXXF -o $ROOT/build/rt/sqlrun $ROOT/tools/sqlrun.cpp -L$ROOT/build/rt -lduckdb -Wl,-rpath,'$ORIGIN' -lpthread -ldl
fi
OUT=$ROOT/oracle/_ref/exon.duckdb_extension
if [ ! -f $OUT ] || [ $ROOT/oracle/ref_loader.cpp -nt $OUT ]; then
    g++ $CXXF -shared -I$REF/exon/include -include duckdb/common/arrow/arrow.hpp \
        $REF/exon/src/exon/arrow_table_function/module.cpp \
        $REF/exon/src/exon/sequence_functions/module.cpp \
        $REF/exon/src/exon/fastq_functions/module.cpp \
        $ROOT/oracle/ref_loader.cpp \
        -L$ROOT/build/rt -lduckdb -L$ROOT/exon_duckdb_b200 -lexon_b200 \
        -Wl,-rpath,'$ORIGIN/../../build/rt' -Wl,-rpath,'$ORIGIN/../../exon_duckdb_b200' -o $OUT
fi
echo "built: $ROOT/build/rt/libduckdb.so $ROOT/build/rt/sqlrun $OUT"
