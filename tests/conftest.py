import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def goldens():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)


# Tests written after the round's last GPU run (the session that added them had no GPU minutes left; their CPU variants are
# green, the device code they drive is the GPU-verified build byte for byte - tools/sass_same_as.sh).  They run LAST, so that a
# surprise in one of them on the B200 cannot hide the verified suite behind `pytest -x`.
_ADDED_AFTER_THE_LAST_GPU_RUN = (        # in the order they run: plain parity first, the multi-threaded ones at the very end
    "test_reference_block8192.py", "test_patch_files.py", "test_getters_reject_null_result_pointers",
    "test_irbank_names_stay_valid_between_adds", "test_edit_between_single_chain_and_batch_calls", "test_batch_files.py",
    "test_spatializer_setters_from_a_second_thread_are_never_lost", "test_tuner_analyze_from_a_second_thread_during_processing",
    "test_one_thread_per_channel_like_the_reference_fan_out",
)


def pytest_collection_modifyitems(config, items):
    def rank(item):
        return next((k for k, name in enumerate(_ADDED_AFTER_THE_LAST_GPU_RUN) if name in item.nodeid), -1)
    late = sorted((i for i in items if rank(i) >= 0), key=rank)          # stable: the file's own order within a rank
    if late:
        items[:] = [i for i in items if rank(i) < 0] + late
