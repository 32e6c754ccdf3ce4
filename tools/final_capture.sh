#!/bin/bash
# tools/final_capture.sh: the end-of-round evidence in one gpurun call (tests, bench lines of both arms, launch list,
# whole-step / SoilCO2 / sweep timings) -> gpurun_out/final/
mkdir -p gpurun_out/final; O=gpurun_out/final
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/pytest_gpu.txt
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --impl reference > $O/bench_reference_arm.json 2>> $O/bench.err
python tools/time_resident_step.py > $O/resident_step.txt 2>&1
python tools/time_soilco2.py > $O/soilco2.txt 2>&1
python tools/sweep.py > $O/sweep.txt 2> $O/sweep.err
python tools/time_hooks.py > $O/hooks.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
tail -2 $O/pytest_gpu.txt; cat $O/bench.json | cut -c1-400; cat $O/resident_step.txt | tail -2; cat $O/soilco2.txt; tail -1 $O/smoke.txt
