#!/bin/bash
# Evidence capture on the GPU box (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash scripts/capture_profiles.sh r01'
# Writes gpurun_out/{bench,ref,launches,prof_*}_<tag>.*; copy the summaries into profiles/ here.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py > $out/bench_${tag}_n1.json 2> $out/bench_${tag}_n1.err
python bench.py --impl reference --steps 10 --warmup 2 > $out/bench_${tag}_ref.json 2> $out/bench_${tag}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-others > $out/ncu_launch_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_assemble_shell_ -s 3 -c 1 \
    -o $out/prof_${tag}_assemble python bench.py --steps 2 --warmup 3 --no-cpu --no-others --no-unstructured > $out/ncu_a_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shell_forces -s 3 -c 1 \
    -o $out/prof_${tag}_forces python bench.py --steps 2 --warmup 3 --no-cpu --no-others --no-unstructured > $out/ncu_f_${tag}.log 2>&1
cat $out/bench_${tag}_n1.json $out/bench_${tag}_ref.json
# frame kernels on a 100^3-joint lattice (2.97 M frames): launch list + full captures
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_${tag}_lattice.csv \
    python scripts/quick_time_lattice.py 100 > $out/ncu_launch_lattice_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame_forces -s 2 -c 1 \
    -o $out/prof_${tag}_frame_forces python scripts/quick_time_lattice.py 100 > $out/ncu_ff_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiles -s 2 -c 1 \
    -o $out/prof_${tag}_frame_tiles python scripts/quick_time_lattice.py 100 > $out/ncu_ft_${tag}.log 2>&1
