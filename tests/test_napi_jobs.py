"""The JS-reachable boundary without Node: napi/apo_napi.c must compile against the N-API prototypes
(vendored stub of the stable C ABI) with -Wall -Wextra -Werror, and the job core it marshals into (napi/apo_jobs.c:
validation + per-handle FIFO) is driven by plain C tests — CPU for the ordering / validation logic, GPU for the
re-entrancy rule through the real C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "napi")]


def test_addon_compiles_against_the_napi_prototypes():
    for src in ("apo_napi.c", "apo_jobs.c"):
        subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", *INC, os.path.join(ROOT, "napi", src)])


def test_addon_exports_cover_the_resident_path_and_comm():
    src = open(os.path.join(ROOT, "napi", "apo_napi.c")).read()
    for name in ("create", "lastCreateError", "dimsUpload", "rolloutsUpload", "corpusUpload", "corpusUploadJson", "scoreResident", "score",
                 "scoreHostRecords", "encodeTuples", "scoreHostTuples", "rewardBatch", "commUniqueId", "commInit", "allocPinned", "recordsFromJson"):
        assert f'{{"{name}", NULL,' in src, name
    assert "napi_throw_error" not in src                      # never throws into the caller (TCS:438)
    ts = open(os.path.join(ROOT, "ts", "apoScoringMainService.ts")).read()
    for name in ("dimsUpload", "corpusUpload", "scoreResident", "scoreHostRecords", "encodeTuples", "scoreHostTuples", "rewardBatch"):
        assert name in ts, name


def _build(apo, tmp_path, name):
    apo.build_library()
    exe = str(tmp_path / name)
    lib_dir = os.path.dirname(apo.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", *INC, os.path.join(ROOT, "tests", "c", name + ".c"),
                           os.path.join(ROOT, "napi", "apo_jobs.c"), "-L", lib_dir, "-lapo_b200", f"-Wl,-rpath,{lib_dir}", "-lpthread", "-lm", "-o", exe])
    return exe


def test_job_fifo_order_and_validation(apo, tmp_path):
    r = subprocess.run([_build(apo, tmp_path, "jobs_order_test")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_overlapping_js_calls_on_one_handle_equal_the_sequential_run(apo, tmp_path):
    r = subprocess.run([_build(apo, tmp_path, "jobs_gpu_test")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("sequential run"), r.stdout + r.stderr
