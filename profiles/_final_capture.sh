set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 300 gpurun_out/bench_r01_final.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err; tail -c 300 gpurun_out/bench_r01_ref.err
timeout 300 python profiles/_step_only.py 3 > gpurun_out/plain_step.log 2>&1; tail -1 gpurun_out/plain_step.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_conv|k_cell|k_scan|k_gather|k_patch|k_compact|k_avgpool" -s 46 -c 23 -o gpurun_out/r01_step_v2 python profiles/_step_only.py 3 > gpurun_out/ncu_step2.log 2>&1; tail -2 gpurun_out/ncu_step2.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_b.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch2.log 2>&1; tail -1 gpurun_out/ncu_launch2.log
