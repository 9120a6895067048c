set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > $O/r4_gpu_tests.log; cat $O/r4_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r4_smoke.log 2>&1; tail -1 $O/r4_smoke.log
timeout 300 python bench.py > $O/r4_bench.json 2> $O/r4_bench.err; cut -c1-200 $O/r4_bench.json
timeout 300 python bench.py --engines 1 > $O/r4_bench_e1.json 2> /dev/null
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r4_bench_ref.json 2> /dev/null; cut -c1-200 $O/r4_bench_ref.json
for w in detect stabilize obb; do timeout 300 python bench.py --workload $w > $O/r4_bench_$w.json 2> /dev/null; done
timeout 300 python bench.py --ingest nv12 > $O/r4_bench_nv12.json 2> /dev/null
timeout 300 python bench.py --workload flight --seconds 20 --steps 20 > $O/r4_flight.json 2> /dev/null; cut -c1-200 $O/r4_flight.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r4_launches.csv python tools/profile_step.py 1 > $O/r4_prof.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/r4_full python tools/profile_step.py 1 > $O/r4_ncu_full.log 2>&1
timeout 300 ncu -i /tmp/r4_full.ncu-rep --page raw --csv > $O/r4_step_full_raw.csv 2>/dev/null; ls -la $O/r4_step_full_raw.csv
GT_TUNE=1 timeout 400 python tools/profile_step.py 1 > $O/r4_tune.out 2> $O/r4_tune.log; grep -c "gt tune" $O/r4_tune.log
timeout 600 python tools/bench_sweeps.py > $O/r4_sweeps.jsonl 2> $O/r4_sweeps.err; wc -l $O/r4_sweeps.jsonl
timeout 600 python tools/bench_registration.py > $O/r4_registration.jsonl 2> $O/r4_registration.err; cat $O/r4_registration.jsonl | cut -c1-400
