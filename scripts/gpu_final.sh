#!/bin/bash
# Final check of a round on one GPU: GPU tests, the bench line, smoke.
tag=${1:-vX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 600 gpurun_out/bench_$tag.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
