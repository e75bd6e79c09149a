mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra"
# launch list of the default bench command (short): shares of the step
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra > /dev/null 2>&1
# --set full of the benchmarked build, one launch each
ncu --set full --clock-control none --import-source on -k regex:rowPipeKernel -s 3 -c 1 -f -o gpurun_out/prof_r02_le $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweepKernel -s 3 -c 1 -f -o gpurun_out/prof_r02_vm $B --workload boxgen200x100x100_c3d8_vonmises > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweepKernel -s 3 -c 1 -f -o gpurun_out/prof_r02_nh $B --workload boxgen100_c3d8tl_neohookewa > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
