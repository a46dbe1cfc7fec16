#!/bin/bash
# round-end measurements on one GPU: headline line, reference arm, every other BASELINE config at its named size,
# launch list of the headline command (ncu, per-launch durations)
set -x
python bench.py --steps 10 --warmup 3 2>gpurun_out/r02_n1.err > gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/r02_n1.err > gpurun_out/r02_bench_reference.json
for c in C1 C2k1 C2k2 C2k3 C4 C5; do
  python bench.py --config $c --steps 2 --warmup 3 2>gpurun_out/r02_$c.err > gpurun_out/r02_bench_$c.json || tail -5 gpurun_out/r02_$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ghb:: -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_bench_under_ncu.json 2>gpurun_out/r02_ncu.err
ls -la gpurun_out/r02_*
