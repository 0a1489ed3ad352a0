mkdir -p gpurun_out
rm -f gpurun_out/prof_r2_step.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 256 > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none -s 12 -c 12 -f -o gpurun_out/prof_r2_step python tools/profile_step.py 2 256 > gpurun_out/prof_full.log 2>&1
ls -la gpurun_out/prof_r2_step.ncu-rep gpurun_out/launches.csv
