#!/bin/bash
D=gpurun_out/s69; mkdir -p $D
timeout 1200 python -m pytest tests -m gpu -x -q > $D/gputest.txt 2>&1; echo "pytest exit $?" | tee -a $D/gputest.txt; tail -3 $D/gputest.txt
python __graft_entry__.py smoke > $D/smoke.txt 2>&1; tail -1 $D/smoke.txt
