#!/bin/bash
# compute-sanitizer over the kernels of the library (GPU box): memcheck and racecheck on the reference fixtures and the
# edge-case tests of both formats, the fused COUNT flavour, the table paths (fused column split, gathers), the scalar
# kernels and the native reader; with >= 2 GPUs, one step of the NVLink peer-memory exchange as well.
# The kernels hand-roll mbarrier / TMA pipelines, decoupled look-backs and system-scope flag protocols: this is the cheap
# evidence that they do not read or write out of bounds and carry no shared-memory hazard.
# usage: bash scripts/gpu_sanitize.sh [outdir]      (logs + a summary; copy the summary to profiles/)
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
SEL=${SEL:-'test_reference_fixtures or test_edge_cases or test_tile_edge_sweep or test_fused_scan_filter_rejects_malformed or test_combined_predicates_and_projection or test_map_fused_into_the_gather or test_gc_content_column or test_quality_score_string_to_list or test_reader_reference_queries or test_computed_columns_fasta or test_string_t_entries or test_fastq_format_matches_the_oracle or test_fasta_format_matches_the_oracle or test_format_is_the_inverse_of_the_scan or test_a_short_output_buffer or test_random_records or test_kernel_inflates_every_block_type or test_kernel_reports_the_first_corrupt_member or test_reader_on_bgzf_fastq_and_fasta or test_rows_of_a_corrupt_bgzf_file or test_second_scan_reads_the_registered_page_cache'}
FILES=${FILES:-"tests/test_gpu_fastq.py tests/test_gpu_fasta.py tests/test_gpu_scalar_reader.py tests/test_gpu_reader2.py tests/test_gpu_writer.py tests/test_inflate.py"}
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --target-processes all --log-file "$OUT/$tool.%p.log" \
      python -m pytest $FILES -m gpu -x -q -k "$SEL" -p no:cacheprovider > "$OUT/$tool.pytest.txt" 2>&1
  echo "== $tool: pytest exit $?" >> "$OUT/summary.txt"
  tail -2 "$OUT/$tool.pytest.txt" >> "$OUT/summary.txt"
  grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|Error:\|Hazard" "$OUT"/$tool.*.log | sort | uniq -c | head -20 >> "$OUT/summary.txt"
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 900 compute-sanitizer --tool memcheck --target-processes all --log-file "$OUT/memcheck_n2.%p.log" \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1 --warmup 3 --reads 200000 --no-e2e --no-cpu --no-paths --no-c5 > "$OUT/memcheck_n2.txt" 2>&1
  echo "== memcheck, N=2 exchange step: exit $?" >> "$OUT/summary.txt"
  grep -h "ERROR SUMMARY" "$OUT"/memcheck_n2.*.log | sort | uniq -c >> "$OUT/summary.txt"
fi
cat "$OUT/summary.txt"
