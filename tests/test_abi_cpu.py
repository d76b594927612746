"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every symbol include/lbbnn.h declares, the
ctypes table binds all of them, host-only queries work without a GPU, and compute calls fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import cases as C  # noqa: F401  (sets sys.path)

HEADER = os.path.join(C.ROOT, "include", "lbbnn.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"LBBNN_API\s+[\w\s\*]+?\b(lbbnn_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    import importlib.util
    spec = importlib.util.spec_from_file_location("lbbnn_build", os.path.join(C.ROOT, "bayesian-neural-nets_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()          # no-op when the in-tree .so is current; nvcc cross-compiles without a GPU


def test_every_declared_symbol_is_exported_and_bound(lib_path):
    names = _declared()
    assert len(names) >= 30
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lbbnn.h but not exported by liblbbnn.so"
    from lbbnn import _capi as K
    assert sorted(K.SIGNATURES) == names, set(K.SIGNATURES) ^ set(names)


def test_host_only_queries_work_without_a_gpu(lib_path):
    from lbbnn import _capi as K
    assert K.lib.lbbnn_abi_version() == 2
    assert K.lib.lbbnn_lrt_f32_workspace_bytes(100, 784, 400) > 0
    assert K.lib.lbbnn_lrt_f32_workspace_bytes(0, 784, 400) == 0
    # the fused-step scheduler is host code: every phase of the MNIST stack fits one round of the persistent grid (2 CTAs per SM)
    st = K.Step()
    st.n_layers, st.batch = 3, 100
    off = 0
    for i, (k, n) in enumerate(C.MNIST_SIZES):
        sl = st.layer[i]
        sl.in_features, sl.out_features = k, n
        sl.off_weight_mu, sl.off_weight_rho, sl.off_lambdal = off, off + k * n, off + 2 * k * n
        sl.off_bias_mu, sl.off_bias_rho = off + 3 * k * n, off + 3 * k * n + (n + 3) // 4 * 4
        off += 3 * k * n + 2 * ((n + 3) // 4 * 4)
        sl.priors, sl.var_mode = K.Priors(0, 1, 0.05, 0, 1), 0
    assert K.lib.lbbnn_lrt_step_workspace_bytes(st) > K.lib.lbbnn_lrt_step_raw_floats(st) * 4 > 0
    buf = ctypes.create_string_buffer(2048)
    assert K.lib.lbbnn_lrt_step_describe(st, buf, 2048) == 0
    desc = buf.value.decode()
    items = [int(a) * int(b) for a, b in re.findall(r"F bn=\d+ kc=\d+ items=(\d+)x(\d+)", desc)]
    assert len(items) == 3 and all(0 < i <= 2 * 148 for i in items), desc
    st.batch = 4096
    assert K.lib.lbbnn_lrt_step_workspace_bytes(st) == 0 and b"batch" in K.lib.lbbnn_last_error()


def test_size_queries_of_the_wide_and_variational_dropout_paths(lib_path):
    """Host-only size / launch-count queries added with the bf16 prologue, the fused head kernel and the variational-dropout
    layer: consistent with the shapes they describe, and usable without a GPU."""
    from lbbnn import _capi as K
    lib = K.lib
    # KL partials of the one-pass bf16 prologue: one double per 64x64 tile
    assert lib.lbbnn_lrt_bf16_prologue_workspace_bytes(4096, 4096) >= 64 * 64 * 8
    assert lib.lbbnn_lrt_bf16_prologue_workspace_bytes(0, 10) == 0
    # column-sum partials of the fused head input gradient: [ceil(batch / 64)][2][in] floats
    assert lib.lbbnn_tc_lrt_bwd_input_small_workspace_bytes(8192, 4096) == 128 * 2 * 4096 * 4
    assert lib.lbbnn_tc_lrt_bwd_input_small_workspace_bytes(65, 8) == 2 * 2 * 8 * 4
    # variational dropout: [gE | gS] staging + the split-contraction partials of the largest GEMM
    for b, n, m in ((100, 784, 1200), (100, 1200, 10), (1, 8, 8), (1000, 400, 72)):
        assert lib.lbbnn_vd_workspace_bytes(b, n, m) >= 2 * b * m * 4
        for shape in ((b, m, n), (n, m, b), (b, n, m)):
            assert lib.lbbnn_vd_gemm_launches(*shape) in (1, 2)
    assert lib.lbbnn_vd_gemm_launches(100, 1200, 1200) == 2       # batch-100 forward: contraction split over gridDim.z
    assert lib.lbbnn_vd_gemm_launches(1200, 1200, 100) == 1       # d theta: enough tiles, epilogue fused
    assert lib.lbbnn_vd_gemm_launches(0, 5, 5) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_a_gpu():
    import lbbnn
    layer = lbbnn.BayesianLinear(8, 4)
    with pytest.raises(lbbnn.LbbnnError):
        layer(torch.zeros(2, 8), sample=True)
    with pytest.raises(lbbnn.LbbnnError):
        lbbnn.LRTTrainer(lbbnn.BayesianNetwork((8, 4, 2)), batch_size=4, num_batches=10)
