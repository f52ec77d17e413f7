#!/bin/bash
# Everything profiles/ holds for one round, in one go.  Run on one B200:
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r2'
# then, back in the container:
#   bash tools/capture_profiles.sh r2 --collect
# gpurun brings back at most 64 MiB and every report embeds the module: reports travel gzipped.
# Numbers printed by runs under ncu are never benchmark values; the bench lines come from the un-profiled runs
# at the top.
TAG=${1:-r2}
OUT=gpurun_out
LEAN="--no-e2e --no-sparse --no-gpu-eager --no-cpu-baseline --no-extras"
NK=16          # kernels matching 'sfr_|decoder_' per iteration of tools/prof_step.py
if [ "$2" == "--collect" ]; then
    for f in $OUT/prof_${TAG}_*.ncu-rep.gz; do gunzip -kf $f; done
    cp $OUT/${TAG}_bench_n1.json $OUT/${TAG}_bench_reference.json $OUT/${TAG}_launches.csv $OUT/${TAG}_sweep_hand17_n1.txt profiles/
    python tools/ncu_summary.py $OUT/prof_${TAG}_kernels.ncu-rep --tag $TAG
    exit 0
fi
mkdir -p $OUT
python bench.py 2> $OUT/${TAG}_bench_n1.err | tail -1 > $OUT/${TAG}_bench_n1.json
python bench.py --impl reference 2> $OUT/${TAG}_bench_reference.err | tail -1 > $OUT/${TAG}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 $LEAN > $OUT/ncu_launches.log 2>&1
# one profiled iteration of every hot kernel (two warm-up iterations skipped)
ncu --set full --clock-control none --import-source on -k regex:'sfr_|decoder_' \
    --launch-skip $((2 * NK)) --launch-count $NK -o $OUT/prof_${TAG}_kernels -f python tools/prof_step.py --iters 3 > $OUT/ncu_kernels.log 2>&1
gzip -f -9 $OUT/prof_${TAG}_*.ncu-rep
python tools/sweep_inference.py > $OUT/${TAG}_sweep_hand17_n1.txt 2> $OUT/sweep.err
tail -c 400 $OUT/${TAG}_bench_n1.json; echo
tail -n 3 $OUT/ncu_kernels.log
ls -la $OUT/prof_${TAG}_*
