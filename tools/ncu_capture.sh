# ncu captures for profiles/: launch lists + one --set full capture per trunk kernel (run under gpurun).
# bench.py's default context has 4 slots -> the 128-channel run profiles trunk_duo_kernel; NSB_TRUNK128=classic
# profiles trunk_fused_kernel<128> (the kernel of one-slot contexts).
CMDDUO="python bench.py --steps 40 --warmup 20 --no-cpu-baseline --no-selfplay --no-latency-leg --small-pool"
CMD256="python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-selfplay --no-latency-leg --small-pool --channels 256 --blocks 20 --batch 512"
$CMDDUO > gpurun_out/plain128.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_128.csv $CMDDUO > gpurun_out/ncu_a.log 2>&1
$CMDDUO > gpurun_out/plain128.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:trunk_duo -s 30 -c 1 -o gpurun_out/prof_duo128_r1 $CMDDUO > gpurun_out/ncu_b.log 2>&1
export NSB_TRUNK128=classic
$CMDDUO > gpurun_out/plain128c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:trunk_fused -s 30 -c 1 -o gpurun_out/prof_trunk128_r1 $CMDDUO > gpurun_out/ncu_b2.log 2>&1
unset NSB_TRUNK128
$CMD256 > gpurun_out/plain256.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:trunk_pair -s 8 -c 1 -o gpurun_out/prof_pair256_r1 $CMD256 > gpurun_out/ncu_c.log 2>&1
$CMD256 > gpurun_out/plain256.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches_256.csv $CMD256 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 2 gpurun_out/ncu_b.log; tail -n 2 gpurun_out/ncu_c.log
