#!/bin/bash
# One GPU-box pass over everything the round's records come from (tag = $1): GPU test suite, smoke, bench lines, public-API
# numbers, then (each only after its plain run exited 0) the ncu launch list of the bench command and one full capture of
# the integration kernel at bench size.  Numbers printed under ncu are never bench values.
T=${1:-r02_final}
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err; tail -c 300 $O/${T}_bench_c2.json
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err; tail -c 200 $O/${T}_bench_c4.json
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu --parity-samples 0 > $O/${T}_bench_c1.json 2> $O/${T}_bench_c1.err
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu --parity-samples 0 > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
timeout 600 python tools/api_bench.py > $O/${T}_api_bench.jsonl 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-dense --parity-samples 0 > $O/${T}_plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_bench_c2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-dense --parity-samples 0 > $O/${T}_ncu_launch.log 2>&1
timeout 300 python tools/prof_run.py c2bench > $O/${T}_plain2.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:diagblk -s 1 -c 1 -f -o $O/${T}_diagblk_c2 \
    python tools/prof_run.py c2bench > $O/${T}_ncu_c2.log 2>&1
tail -2 $O/${T}_ncu_c2.log
