"""The C-ABI library loads on a CPU-only box, exports every symbol include/apo_b200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "apo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apo_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(apo):
    assert header_symbols() == sorted(apo.ABI_SYMBOLS)


def test_library_exports_every_symbol(apo):
    apo.build_library()
    lib = ctypes.CDLL(apo.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), s
    assert lib.apo_abi_version() == 2


def test_struct_sizes_match_header(apo):
    from importlib import import_module
    eng = import_module("senweaver-ide_b200.engine")
    assert ctypes.sizeof(eng.Pattern) == 40
    assert ctypes.sizeof(eng.DimStat) == 32
    assert ctypes.sizeof(eng.ScoreOpts) == 32
    assert ctypes.sizeof(eng.Timing) == 44
    assert eng.RECORD_DTYPE.itemsize == 32


def test_no_cpu_fallback(apo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    apo.build_library()
    with pytest.raises(apo.ApoError) as ei:
        apo.Engine(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "senweaver-ide_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "apo_oracle" not in txt, f


def test_ts_codec_offsets_match_the_c_structs(apo):
    """ts/traceRecordCodec.ts hard-codes the byte offsets of apo_corpus_report; pin them to the ctypes mirror."""
    from importlib import import_module
    eng = import_module("senweaver-ide_b200.engine")
    R, D, P = eng.CorpusReport, eng.DimStat, eng.Pattern
    assert ctypes.sizeof(R) == 784
    off = {f: getattr(R, f).offset for f, _ in R._fields_}
    assert (off["total"], off["good"], off["bad"], off["none"], off["goodRate"], off["byMode"], off["byModeGoodRate"]) == (0, 8, 16, 24, 32, 40, 160)
    assert (off["withReward"], off["avgReward"], off["dim"], off["pat"], off["toolCalls"], off["toolSucc"], off["toolFail"], off["toolSuccessRate"]) == \
           (200, 216, 224, 512, 752, 760, 768, 776)
    assert (ctypes.sizeof(D), D.count.offset, D.avg.offset, D.low_flag.offset, D.low_severity.offset, D.sugg_flag.offset, D.sugg_priority.offset) == (32, 8, 16, 24, 25, 26, 27)
    assert (ctypes.sizeof(P), P.flag.offset, P.severity.offset, P.examples.offset) == (40, 8, 9, 16)
    src = open(os.path.join(ROOT, "ts", "traceRecordCodec.ts")).read()
    for token in ("224 + 32 * i", "512 + 40 * p", "160 + 8 * m", "40 + 24 * m", "u64(v, 752)", "v.getFloat64(776, true)", "784"):
        assert token in src, token


def _build_c_demo(apo, tmp_path):
    import subprocess
    apo.build_library()
    exe = str(tmp_path / "score_demo")
    lib_dir = os.path.dirname(apo.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "score_demo.c"),
                           "-L", lib_dir, "-lapo_b200", f"-Wl,-rpath,{lib_dir}", "-lm", "-o", exe])
    return exe


def test_plain_c_client_links_against_the_abi(apo, tmp_path):
    """examples/score_demo.c: pure C11, linked against libapo_b200.so; without a GPU it must fail loudly (exit 2)."""
    import subprocess
    import torch
    exe = _build_c_demo(apo, tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
        assert "top-4:" in r.stdout and "K > C -> rc=-1" in r.stdout
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stdout
