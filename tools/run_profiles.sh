#!/bin/bash
# Run ON THE GPU BOX (gpurun -- 'bash tools/run_profiles.sh'): the captures tools/make_profiles.py turns into profiles/.
# ncu per-launch times are cold-cache and serialised; bench numbers come from the separate, un-profiled bench runs.
mkdir -p gpurun_out
rm -f gpurun_out/prof_r2_step.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 256 > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none -s 12 -c 12 -f -o gpurun_out/prof_r2_step python tools/profile_step.py 2 256 > gpurun_out/prof_full.log 2>&1
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv > gpurun_out/clocks_after_profiles.csv
ls -la gpurun_out/prof_r2_step.ncu-rep gpurun_out/launches.csv
tail -c 600 gpurun_out/bench.log
