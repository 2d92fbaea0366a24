#!/bin/bash
D=gpurun_out/s56; mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_reader2.py tests/test_inflate.py -m gpu -x -q 2>&1 | tail -3
python scripts/probe_register_sequence.py 2>&1 | tail -8 | tee $D/register_sequence.txt
EXON_B200_REGISTER_PIECE_MB=16 python scripts/probe_register_sequence.py 2>&1 | tail -8 | tee $D/register_sequence_16.txt
