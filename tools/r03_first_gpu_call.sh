#!/bin/bash
# First gpurun call of a next round: what the last session of round 2 could not run on a B200 (its GPU minutes were spent) -
# the -m gpu variants of the tests it added (host-side changes only: profiles/r02_host_sanitizers.txt), then the whole GPU
# suite, the default bench and the launch list.       gpurun --timeout 2400 -- 'bash tools/r03_first_gpu_call.sh'
mkdir -p gpurun_out
T=${TAG:-r03a}
timeout 600 python -m pytest -q -m gpu -x \
    tests/test_edge_cases.py::test_edit_between_single_chain_and_batch_calls \
    tests/test_edge_cases.py::test_one_thread_per_channel_like_the_reference_fan_out \
    tests/test_next_rows.py::test_spatializer_setters_from_a_second_thread_are_never_lost \
    tests/test_next_rows.py::test_tuner_analyze_from_a_second_thread_during_processing \
    tests/test_next_rows.py::test_getters_reject_null_result_pointers \
    tests/test_next_rows.py::test_irbank_names_stay_valid_between_adds \
    tests/test_patch_files.py tests/test_reference_block8192.py tests/test_batch_files.py 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_new_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -c 600 gpurun_out/${T}_bench_n1.json
TAG=$T bash tools/r02_profiles.sh
