"""CPU-side checks: the C-ABI library loads, exports every symbol include/gccvae.h declares, and the
host layer fails loudly (no fallback) when there is no GPU."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="gccvae.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gccvae_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_exports_header():
    import gccvae_b200._lib as L
    assert os.path.exists(L.LIB_PATH), "run `python __graft_entry__.py build` first"
    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libgccvae.so does not export " + name
    assert set(declared) == set(L.SIGNATURES), (set(declared) ^ set(L.SIGNATURES))
    assert lib.gccvae_abi_version() == 1
    # the development aids live in their own header, outside the boundary
    debug = _declared_symbols("gccvae_debug.h")
    assert set(debug) == set(L.DEBUG_SIGNATURES) and all(n.startswith("gccvae_debug_") for n in debug)
    assert not any(n.startswith("gccvae_debug_") for n in declared)
    for name in debug:
        assert hasattr(lib, name)


def test_param_store_layout_matches_reference_counts():
    from gccvae_b200.params import ParamStore, keras_default_init
    s = ParamStore("cpu")
    assert s.numel(True) == 1007901 and s.numel(False) == 1007577
    count = lambda pre: sum(n for k, (_, n, _) in s.offsets.items() if k.startswith(pre))
    assert (count("enc."), count("dec."), count("cls."), count("prior."), count("mu")) == (729690, 276249, 342, 1296, 324)
    assert all(off % 4 == 0 for off, _, _ in s.offsets.values())
    assert list(s.offsets)[-1] == "mu" and s.n_without_mu == s.offsets["mu"][0]
    keras_default_init(s, 0)
    assert float(s.view("prior.scale_true").min()) == 1.0 and float(s.view("prior.loc_true").abs().max()) == 0.0
    assert float(s.view("enc.conv1.b").abs().max()) == 0.0
    lim = (6.0 / (16 * 3 + 16 * 32)) ** 0.5
    assert float(s.view("enc.conv1.w").abs().max()) <= lim


def test_oracle_and_store_share_the_parameter_table():
    import gccvae_oracle as O
    from gccvae_b200.params import param_specs
    assert [(k, tuple(v)) for k, v in O.PARAM_SHAPES] == [(k, tuple(v)) for k, v in param_specs()[:-1]]


@pytest.mark.parametrize("frac", ["0.0", "0.1", "0.2", "0.5", "1.0"])
def test_gating_matrix_loading_is_bit_exact(golden_dir, frac):
    import gccvae_b200 as G
    want = np.load(os.path.join(golden_dir, "data", "gating_matrix_{}.npy".format(frac)))
    got = G.load_gating_matrix(os.path.join(golden_dir, "data"), frac)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    rdr = G.GatingMatrixReader(os.path.join(golden_dir, "data"), float(frac) if frac != "0.0" else 0.0)
    assert np.array_equal(rdr.init_gating_prob, want)


def test_missing_gating_matrix_raises(tmp_path):
    import gccvae_b200 as G
    with pytest.raises(FileNotFoundError):
        G.load_gating_matrix(str(tmp_path), 0.3)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    import gccvae_b200 as G
    cfg = dict(gate_type="fixed", gate_subtype="one-one", lr=1e-4, gating_init_temp=0.3)
    with pytest.raises(G.GccvaeError, match="no CPU fallback"):
        G.Learner((64, 64, 3), 45, 18, 18, 10, 1.0, cfg)
    with pytest.raises(G.GccvaeError):
        G.img_log_likelihood(torch.zeros(1, 64, 64, 3), torch.zeros(1, 64, 64, 3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "semi-supervised-gated-lt-vae_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "gccvae_oracle" not in src and "oracle/" not in src, f
