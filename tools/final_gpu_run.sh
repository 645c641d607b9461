#!/bin/bash
# One GPU-box pass over everything the round's records come from (tag = $1): GPU test suite, smoke, bench lines,
# then (each only after its plain run exited 0) the ncu launch list of the bench command and one full capture of
# the integration kernel.  Numbers printed under ncu are never bench values.
T=${1:-s51}
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
timeout 600 python bench.py > $O/${T}_bench_c2.log 2> $O/${T}_bench_c2.err; tail -c 400 $O/${T}_bench_c2.log
timeout 300 python bench.py --workload c4 --no-cpu > $O/${T}_bench_c4.log 2> $O/${T}_bench_c4.err; tail -c 300 $O/${T}_bench_c4.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-dense > $O/${T}_plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_bench_c2_launches_v4.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-dense > $O/${T}_ncu_launch.log 2>&1
timeout 300 python tools/prof_run.py c2 > $O/${T}_plain2.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_diag_kernel -c 1 -f -o $O/r01_diag_c2_cw \
    python tools/prof_run.py c2 > $O/${T}_ncu_c2.log 2>&1
tail -2 $O/${T}_ncu_c2.log
