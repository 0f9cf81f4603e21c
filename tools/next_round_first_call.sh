#!/bin/bash
# One gpurun command line for the first GPU call of the next round (every call costs >= 1-2 GPU-minutes of box time,
# so batch): the pending device tests (tests/pending), the whole GPU suite, the bench line, the perf matrix.
#   tools/gpurun_retry.sh 2400 "$(cat tools/next_round_first_call.sh | grep -v '^#' | tr '\n' ' ')"
mkdir -p gpurun_out;
python -m pytest tests/pending/gpu_ordered_events.py tests/pending/gpu_sector_planes.py tests/pending/gpu_matrix_values.py tests/pending/gpu_config_sizes.py tests/pending/gpu_tile_taper.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pending_tests.txt;
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/gpu_suite.txt;
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json;
timeout 1200 python tools/perf_matrix.py 2>&1 | tee gpurun_out/perf_matrix.txt | tail -30;
