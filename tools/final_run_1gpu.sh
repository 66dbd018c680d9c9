#!/bin/bash
# Round-end style single-GPU run: smoke, the GPU test suite, both bench arms, then the ncu launch list and the DRAM-traffic
# pass of the SAME bench command (numbers printed under ncu are never bench values).  Outputs under gpurun_out/final/.
set -u
O=gpurun_out/final
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; tail -c 300 $O/bench_reference_arm.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -c 300 $O/bench_1gpu.json
CMD="python bench.py --cycle-only --no-cpu-baseline --no-other-configs --steps 2 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_config_D.csv $CMD > $O/ncu_launches.log 2>&1
python tools/launch_summary.py $O/launches_config_D.csv 30 > $O/launch_summary_config_D.txt 2>&1; head -8 $O/launch_summary_config_D.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:smooth3d32c --launch-skip 48 -c 16 --csv --log-file $O/ncu_config_D_smoother_dram.csv $CMD > $O/ncu_dram.log 2>&1
tail -3 $O/ncu_config_D_smoother_dram.csv | cut -c1-300
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
