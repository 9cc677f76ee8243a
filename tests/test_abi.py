"""CPU: the C-ABI library loads and exports exactly what include/healswin_b200.h declares."""
import ctypes
import os
import re

from heal_swin_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "healswin_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hs_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 10
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in heal_swin_b200/_lib.py"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"


def test_version_and_error_string():
    assert _lib.lib.hs_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_library_has_no_torch_or_libcuda_link_dependency():
    import subprocess

    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libc10" not in out


def test_shape_predicates_are_host_only():
    """The *_supported predicates are pure host functions (no device needed): the shapes of BASELINE configs[1] are
    covered, shapes outside the kernels' tiling are not."""
    lib = _lib.lib
    T0 = 8 * 196608
    # weight gradient: qkv / fc1 (bias fused), fc2 (N < K: no fused bias), stage 3 (min(N, K) = 768: two column halves)
    assert lib.hs_linear_wgrad_supported(T0, 288, 96) == 2 and lib.hs_linear_wgrad_supported(T0, 384, 96) == 2
    assert lib.hs_linear_wgrad_supported(T0, 96, 384) == 1
    assert lib.hs_linear_wgrad_supported(T0 // 64, 2304, 768) == 1 and lib.hs_linear_wgrad_supported(T0, 10, 96) == 0
    assert lib.hs_linear_wgrad_supported(T0 // 64, 3072, 1536) == 0  # min(N, K) > 1024
    assert lib.hs_linear_wgrad_supported(100, 288, 96) == 0
    # fused MLP backward: stages 0-1
    assert lib.hs_mlp_dgrad_gelu_supported(T0, 96, 384) == 1 and lib.hs_mlp_dgrad_gelu_supported(T0 // 4, 192, 768) == 1
    assert lib.hs_mlp_dgrad_gelu_supported(T0 // 16, 384, 1536) == 0 and lib.hs_mlp_dgrad_gelu_supported(T0, 96, 100) == 0
    # bf16x3 GEMM: feature dimensions multiples of 4 (TMA row pitch); the 10-class head and odd widths are not
    assert lib.hs_gemm3_supported(T0, 288, 96) == 1 and lib.hs_gemm3_supported(T0, 96, 12) == 1
    assert lib.hs_gemm3_supported(T0, 144, 48) == 1 and lib.hs_gemm3_supported(T0 // 64, 3072, 768) == 1
    assert lib.hs_gemm3_supported(T0, 10, 96) == 0 and lib.hs_gemm3_supported(T0, 96, 50) == 0
    assert lib.hs_bias_gelu_supported(T0, 384) == 1 and lib.hs_bias_gelu_supported(T0, 6144) == 0
    assert lib.hs_bias_gelu_supported(T0, 50) == 0
    # fused decoder tail: C in {32, 64, 96, 128}, up to 16 output channels
    assert lib.hs_ln_head_supported(4 * T0, 96, 10) == 1 and lib.hs_ln_head_supported(4 * T0, 96, 1) == 1
    assert lib.hs_ln_head_supported(4 * T0, 128, 1) == 1 and lib.hs_ln_head_supported(4 * T0, 96, 17) == 0
    assert lib.hs_ln_head_supported(4 * T0, 160, 1) == 0


def test_argument_validation_happens_before_any_device_work():
    """Bad arguments are rejected on the host with HS_ERR_ARG / HS_ERR_UNSUPPORTED and a message (no GPU needed)."""
    import ctypes as C

    lib = _lib.lib
    null = None
    one = C.c_void_p(16)  # a non-null, 16-byte aligned dummy: must never be dereferenced by the checks below
    rc = lib.hs_ln_head_fwd(null, one, one, one, null, one, one, one, 8, 4, 96, 10, C.c_float(1e-5), null)
    assert rc == 1 and "null" in _lib.last_error()
    rc = lib.hs_ln_head_fwd(one, one, one, one, null, one, one, one, 10, 4, 96, 10, C.c_float(1e-5), null)
    assert rc == 1 and "rows_per_sample" in _lib.last_error()
    rc = lib.hs_ln_head_fwd(one, one, one, one, null, one, one, one, 8, 4, 48, 10, C.c_float(1e-5), null)
    assert rc == 3 and "not covered" in _lib.last_error()
    rc = lib.hs_mlp_dgrad_gelu(one, one, one, null, C.c_float(1.5), 0, one, 8192, 96, 384, 0, null)
    assert rc == 1 and "drop" in _lib.last_error()
    rc = lib.hs_mlp_dgrad_gelu(one, one, one, null, C.c_float(0.0), 0, one, 8192, 384, 1536, 0, null)
    assert rc == 3 and "not covered" in _lib.last_error()
    rc = lib.hs_gemm3(null, one, null, null, one, null, null, 8, 32, 32, 0, 0, C.c_float(0.0), 0, null)
    assert rc == 1 and "bad arguments" in _lib.last_error()
    rc = lib.hs_gemm3(one, one, null, null, one, null, null, 8, 32, 32, 1, 0, C.c_float(0.0), 0, null)
    assert rc == 1 and "aux" in _lib.last_error()
    rc = lib.hs_gemm3(one, one, null, null, one, null, null, 8, 32, 32, 2, 0, C.c_float(0.0), 0, null)
    assert rc == 1 and "second output" in _lib.last_error()
    rc = lib.hs_gemm3(one, one, null, null, one, null, null, 8, 30, 32, 0, 0, C.c_float(0.0), 0, null)
    assert rc == 3 and "not covered" in _lib.last_error()
    rc = lib.hs_weight_split(null, 4, 4, 4, 0, 0, one, null)
    assert rc == 1


def test_every_entry_point_is_documented_for_the_integrator():
    """INTEGRATION.md's table names every exported entry point together with the reference code it replaces."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_functions() if n not in text]
    assert not missing, f"not mentioned in INTEGRATION.md: {missing}"
