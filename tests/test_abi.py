"""The C-ABI library loads and exports every symbol include/avsync.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "avsync.h")


def declared_symbols(experiments: bool = False):
    src = open(HEADER).read()
    exp_blocks = re.findall(r"#ifdef AVS_EXPERIMENTS(.*?)#endif", src, flags=re.S)
    if experiments:
        src = "\n".join(exp_blocks)
    else:
        src = re.sub(r"#ifdef AVS_EXPERIMENTS.*?#endif", "", src, flags=re.S)
    return sorted(set(re.findall(r"AVS_API\s+[\w\s\*]+?\b(avs_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def native():
    import avsync_b200
    N = avsync_b200._native
    if not os.path.isfile(N.LIB_PATH):
        N.build()
    return N


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("avs_mfcc_stats_sweep", "avs_stcnn_forward", "avs_bigru_forward", "avs_sweep_score",
                 "avs_ctc_greedy", "avs_sweep_run", "avs_sweep_run_host", "avs_sweep_run_u8", "avs_sweep_run_host_u8",
                 "avs_stcnn_forward_u8", "avs_source_hash", "avs_version", "avs_last_error_string"):
        assert must in syms
    assert len(syms) >= 27


def test_library_exports_every_declared_symbol(native):
    L = ctypes.CDLL(native.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in avsync.h but not exported"


def test_binding_covers_every_declared_symbol(native):
    assert sorted(native.SIGNATURES) == declared_symbols()
    assert sorted(native.EXPERIMENT_SIGNATURES) == declared_symbols(experiments=True)
    assert native.lib().avs_version() == 200


def test_product_library_carries_no_experiment_switches(native):
    """The experiment switches (avs_debug_set, AVS_* environment knobs) live in the tools build only
    (make EXPERIMENTS=1 -> libavsync_b200_exp.so): the product binary neither exports the switch nor reads the
    environment, and its baked-in source hash matches the tree it was loaded from."""
    import subprocess
    L = ctypes.CDLL(native.LIB_PATH)
    for s in declared_symbols(experiments=True):
        assert not hasattr(L, s), f"{s} must not be exported by the product library"
    undefined = subprocess.run(["nm", "-D", "--undefined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "getenv" not in undefined
    assert native.lib().avs_source_hash().decode() == native.source_hash()


def test_no_cpu_fallback_without_cuda(native):
    import torch
    import avsync_b200 as A
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    net = A.LipNet(39).eval()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 75, 50, 100))
    with pytest.raises(RuntimeError):
        A.compute_audio_stats(__import__("numpy").zeros(48000, dtype="float32"), 16000, 20)
    with pytest.raises(RuntimeError):
        A.ctc_greedy_decode(torch.zeros(1, 75, 39))


def test_argument_validation_returns_error_codes_without_a_device(native):
    """Bad arguments are rejected with AVS_EINVAL (-1) and a message before any CUDA call is made."""
    import ctypes
    L = native.lib()
    arr = (ctypes.c_int32 * 3)(0, 640, -640)
    h = ctypes.c_void_p()
    assert L.avs_mfcc_plan_create(48000, 16000, 0, arr, 3, ctypes.byref(h)) == -1          # n_mfcc out of range
    assert b"n_mfcc" in L.avs_last_error_string()
    assert L.avs_mfcc_plan_create(48000, 16000, 41, arr, 3, ctypes.byref(h)) == -1
    assert L.avs_mfcc_plan_create(0, 16000, 20, arr, 3, ctypes.byref(h)) == -1             # empty signal
    assert L.avs_mfcc_plan_create(48000, 16000, 20, None, 3, ctypes.byref(h)) == -1        # null shifts
    nf, nu = ctypes.c_int(), ctypes.c_int()
    assert L.avs_mfcc_plan_describe(48000, 16000, arr, 0, ctypes.byref(nf), ctypes.byref(nu), None, None) == -1
    assert L.avs_sweep_score(None, None, 1, 41, 13824, 40, None, None, None, None, 512, None, None, None, 0, None) == -1
    assert L.avs_ctc_greedy(None, 1, 75, 39, 0, None, None, None) == -1
    assert L.avs_preproc_create(288, 360, 2, ctypes.byref(h)) == -1                        # channels must be 1 or 3
    with pytest.raises(ValueError):
        import avsync_b200 as A
        A.LipNet(39, precision="fp8")
