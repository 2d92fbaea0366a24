#!/bin/bash
# sanitizer over the new code (inflate kernel, BGZF reader paths, registered page cache) + the registration sequence probe
export FILES="tests/test_inflate.py tests/test_gpu_reader2.py"
export SEL="test_kernel_inflates_every_block_type or test_kernel_reports_the_first_corrupt_member or test_reader_on_bgzf_fastq_and_fasta or test_rows_of_a_corrupt_bgzf_file or test_second_scan_reads_the_registered_page_cache or test_reader_reports_a_corrupt_bgzf_member"
bash scripts/gpu_sanitize.sh gpurun_out/s55 2>&1 | tail -12
python scripts/probe_register_sequence.py 2>&1 | tail -12
