# Round-2 final N=1 records (run under gpurun from the repo root): GPU test suite, both bench arms, the pyramid config,
# one `ncu --set full` capture of a resident step and the launch list of the bench command.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 500 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.err
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -c 300 gpurun_out/r02_bench_reference_arm.err
timeout 400 python tools/run_config.py pyramid > gpurun_out/r02_cfg_pyramid.json 2> gpurun_out/r02_cfg_pyramid.err; tail -c 300 gpurun_out/r02_cfg_pyramid.err
timeout 300 python profiles/_step_only.py 3 > gpurun_out/plain_step.log 2>&1; tail -1 gpurun_out/plain_step.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_conv|k_cell|k_scan|k_gather|k_patch|k_compact|k_avgpool" -s 48 -c 24 -f -o gpurun_out/r02_step_final python profiles/_step_only.py 3 > gpurun_out/ncu_step_final.log 2>&1; tail -2 gpurun_out/ncu_step_final.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_b.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ncu_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_final.log 2>&1; tail -1 gpurun_out/ncu_launch_final.log
