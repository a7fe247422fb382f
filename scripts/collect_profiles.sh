#!/bin/bash
# copy the summaries of a capture (scripts/capture_profiles.sh <tag>) from gpurun_out/ into profiles/
tag=$1
o=gpurun_out; p=profiles
cp $o/bench_${tag}_n1.json $p/bench_${tag}_n1.json
cp $o/bench_${tag}_ref.json $p/bench_${tag}_ref.json
cp $o/launches_${tag}.csv $p/${tag}_launches.csv
cp $o/launches_${tag}_lattice.csv $p/${tag}_launches_lattice.csv
for k in assemble forces frame_forces frame_tiles; do
  python scripts/ncu_details.py $o/prof_${tag}_$k.ncu-rep > $p/${tag}_prof_$k.txt
  python scripts/ncu_summary.py $o/prof_${tag}_$k.ncu-rep >> $p/${tag}_prof_$k.txt
done
