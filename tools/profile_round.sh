#!/bin/bash
# tools/profile_round.sh tag : the ncu evidence of a round for the default workload (configs[2], 65 536 records per step =
# two prepare / solve / certify passes of 32 768): launch list + one full capture of each kernel.  Run under gpurun AFTER the
# same command has exited 0 without ncu.
tag=$1
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency"
$cmd > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { tail -5 gpurun_out/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu_launches.log 2>&1
for k in solve factor certify; do
  ncu --set full --import-source on --clock-control none -k regex:qp_$k -s 6 -c 1 -o gpurun_out/${tag}_$k -f $cmd > gpurun_out/${tag}_ncu_$k.log 2>&1
  tail -1 gpurun_out/${tag}_ncu_$k.log
done
