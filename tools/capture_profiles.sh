#!/bin/bash
# Everything profiles/ holds for one round, in one go.  Run on one B200:
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r1'
# then, back in the container:
#   bash tools/capture_profiles.sh r1 --collect
# gpurun brings back at most 64 MiB and every report embeds the 12 MB module: reports travel gzipped.
# Numbers printed by runs under ncu are never benchmark values; the bench lines come from the two
# un-profiled runs at the top.
TAG=${1:-r1}
OUT=gpurun_out
LEAN="--no-e2e --no-sparse --no-gpu-eager --no-cpu-baseline"
if [ "$2" == "--collect" ]; then
    for f in $OUT/prof_${TAG}_*.ncu-rep.gz; do gunzip -kf $f; done
    cp $OUT/${TAG}_bench_n1.json $OUT/${TAG}_bench_reference.json $OUT/${TAG}_launches.csv profiles/
    python tools/ncu_summary.py $OUT/prof_${TAG}_step.ncu-rep --tag $TAG
    python tools/ncu_summary.py $OUT/prof_${TAG}_two.ncu-rep --tag $TAG --append "two-kernel route of the last stage (SURVEY 8d accounting): forward kernel, backward+loss kernel"
    python tools/ncu_summary.py $OUT/prof_${TAG}_lean.ncu-rep --tag $TAG --append "compact-target step: one-pass last stage evaluating the targets from 64-byte taps"
    python tools/ncu_summary.py $OUT/prof_${TAG}_infer.ncu-rep --tag $TAG --joints 21 --append "inference pass (HAND17): test-only SFR + pipelined forward without the heat-map store"
    python tools/profiles_readme.py $TAG
    exit 0
fi
mkdir -p $OUT
python bench.py 2> $OUT/${TAG}_bench_n1.err | tail -1 > $OUT/${TAG}_bench_n1.json
python bench.py --impl reference 2> $OUT/${TAG}_bench_reference.err | tail -1 > $OUT/${TAG}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 $LEAN > $OUT/ncu_launches.log 2>&1
# default step: 2 hot kernels (sfr_build_kernel, decoder_fused_kernel), 3 warm-up steps -> the first timed step
ncu --set full --clock-control none --import-source on -k regex:'sfr_build_kernel|decoder_fused_kernel' \
    --launch-skip 6 -c 2 -o $OUT/prof_${TAG}_step -f python bench.py --steps 2 --warmup 3 $LEAN > $OUT/ncu_step.log 2>&1
# variants timed after the main region: (5 default steps = 5 sfr_build launches, then) the two-kernel route
# (3 warm-up steps x 3 matching kernels), then the compact-target step
ncu --set full --clock-control none --import-source on -k regex:'sfr_build_kernel|decoder_fwd_kernel|decoder_bwd_pipe_kernel' \
    --launch-skip 14 -c 3 -o $OUT/prof_${TAG}_two -f python bench.py --steps 2 --warmup 3 --no-e2e --no-gpu-eager --no-cpu-baseline > $OUT/ncu_two.log 2>&1
# compact-target step (timed last): 5 default steps x 2 matching kernels + 5 sfr_build launches of the
# two-kernel steps + 3 compact warm-up steps x 2
ncu --set full --clock-control none --import-source on -k regex:'sfr_build_kernel|decoder_fused_kernel' \
    --launch-skip 21 -c 2 -o $OUT/prof_${TAG}_lean -f python bench.py --steps 2 --warmup 3 --no-e2e --no-gpu-eager --no-cpu-baseline > $OUT/ncu_lean.log 2>&1
# inference pass of tools/sweep_inference.py
ncu --set full --clock-control none --import-source on -k regex:'sfr_build_kernel|decoder_fwd_pipe_kernel' \
    --launch-skip 4 -c 2 -o $OUT/prof_${TAG}_infer -f python tools/sweep_inference.py --batches 4096 --steps 2 --warmup 1 > $OUT/ncu_infer.log 2>&1
gzip -f -9 $OUT/prof_${TAG}_*.ncu-rep
python tools/sweep_inference.py > $OUT/${TAG}_sweep_hand17_n1.txt 2> $OUT/sweep.err
tail -1 $OUT/${TAG}_bench_n1.json | cut -c1-400
for f in ncu_step ncu_two ncu_lean ncu_infer; do tail -n 2 $OUT/$f.log; done
