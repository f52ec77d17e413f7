#!/bin/bash
# compute-sanitizer over the one-pass last-stage kernel (all 12 instantiations, ~7 items per CTA so every
# ring slot is re-used; also as the inner-stage forward), the six-slot backward, the compact-target comparison,
# fetch + window mode, the raw-frame window table, the optional staged SFR
# builder, the bb loader and smoke().
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh r2'
# Logs land in gpurun_out/<tag>_sanitizer_<tool>.log; copy the summaries into profiles/.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
TESTS="tests/test_gpu_pins.py::test_one_pass_kernel_every_instantiation_many_items_per_cta tests/test_gpu_pins.py::test_lean_one_pass_kernel_equals_the_two_cta_one_pass_kernel tests/test_gpu_decoder.py::test_sparse_targets_equal_dense_targets tests/test_gpu_pins.py::test_n_mean_two_half_batches_sum_to_the_full_batch tests/test_gpu_pins.py::test_inner_stage_forward_loss_through_the_one_pass_kernel_equals_the_direct_kernel tests/test_gpu_decoder.py::test_forward_backward_match_fp64_oracle tests/test_gpu_decoder.py::test_half_precision_conv_outputs_equal_upcast_path tests/test_gpu_feed.py tests/test_gpu_sfr.py::test_raw_frame_window_table_equals_per_tap_arithmetic tests/test_gpu_sfr.py::test_gpu_raw_frames_bitwise_equal_oracle tests/test_gpu_sfr.py::test_gpu_hand17_bb_loader_matches_reference_golden tests/test_gpu_sfr.py::test_gpu_matches_reference_golden tests/test_gpu_sfr.py::test_staged_source_rows_equal_the_direct_gather"
for tool in racecheck memcheck initcheck; do
    timeout 1500 compute-sanitizer --tool $tool --log-file $OUT/${TAG}_sanitizer_${tool}.raw \
        python -m pytest $TESTS -x -q -p no:cacheprovider > $OUT/${TAG}_sanitizer_${tool}.pytest 2>&1
    echo "pytest exit $?" >> $OUT/${TAG}_sanitizer_${tool}.pytest
    { echo "# compute-sanitizer --tool $tool python -m pytest $TESTS"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|Uninitialized|hazard" $OUT/${TAG}_sanitizer_${tool}.raw | sort | uniq -c | head -40; tail -3 $OUT/${TAG}_sanitizer_${tool}.pytest; } > $OUT/${TAG}_sanitizer_${tool}.log
    head -c 200000 $OUT/${TAG}_sanitizer_${tool}.raw > $OUT/${TAG}_sanitizer_${tool}.head; rm -f $OUT/${TAG}_sanitizer_${tool}.raw
done
timeout 600 compute-sanitizer --tool racecheck --log-file $OUT/${TAG}_sanitizer_smoke.raw python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_sanitizer_smoke.out 2>&1
{ echo "# compute-sanitizer --tool racecheck smoke()"; grep -E "SUMMARY|hazard" $OUT/${TAG}_sanitizer_smoke.raw | sort | uniq -c; tail -2 $OUT/${TAG}_sanitizer_smoke.out; } > $OUT/${TAG}_sanitizer_smoke.log
rm -f $OUT/${TAG}_sanitizer_smoke.raw
cat $OUT/${TAG}_sanitizer_*.log
