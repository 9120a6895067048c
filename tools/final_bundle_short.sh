set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > $O/r6_gpu_tests.log; cat $O/r6_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r6_smoke.log 2>&1; tail -1 $O/r6_smoke.log
timeout 300 python bench.py > $O/r6_bench.json 2> $O/r6_bench.err; cut -c1-200 $O/r6_bench.json
timeout 300 python bench.py --engines 1 > $O/r6_bench_e1.json 2> /dev/null
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r6_bench_ref.json 2> /dev/null; cut -c1-200 $O/r6_bench_ref.json
timeout 300 python bench.py --workload stabilize > $O/r6_bench_stabilize.json 2> /dev/null
timeout 300 python bench.py --workload flight --seconds 20 --steps 20 > $O/r6_flight.json 2> /dev/null; cut -c1-200 $O/r6_flight.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r6_launches.csv python tools/profile_step.py 1 > $O/r6_prof.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/r6_full python tools/profile_step.py 1 > $O/r6_ncu_full.log 2>&1
timeout 300 ncu -i /tmp/r6_full.ncu-rep --page raw --csv > $O/r6_step_full_raw.csv 2>/dev/null; ls -la $O/r6_step_full_raw.csv
