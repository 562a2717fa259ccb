"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/qtcnn.h
declares, the ctypes signature table covers them all, the module mirror keeps the reference's constructor
surface / state_dict keys, and the product path refuses to run without a GPU (no CPU fallback)."""
import ctypes

import pytest
import torch

import qtcnn_b200.capi as capi


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    declared = capi.declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"libqtcnn.so does not export {name}"
    assert set(declared) == set(capi._SIGNATURES), set(declared) ^ set(capi._SIGNATURES)
    assert lib.qt_version() >= 100


def test_conv_desc_matches_c_struct_layout():
    d = capi.conv_desc(2, (1, 14, 14), 256, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    assert ctypes.sizeof(capi.ConvDesc) == 16 * 4 + 16 * 8
    assert (d.x_stride[0], d.x_stride[2], d.x_stride[3]) == (14 * 14 * 256, 14 * 256, 256)
    # 3x3/s1: persistent slab kernel -> one BatchNorm partial row per CTA (= number of 256-pixel tiles of the padded map)
    assert capi.lib().qt_conv_stat_rows(d) == (2 * 16 * 16 + 255) // 256
    # strided conv: generic gather kernel -> one row per 128-pixel tile
    d2 = capi.conv_desc(2, (1, 14, 14), 256, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    assert capi.lib().qt_conv_stat_rows(d2) == (2 * 7 * 7 + 127) // 128
    bad = capi.conv_desc(2, (1, 14, 14), 250, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    assert capi.lib().qt_conv_stat_rows(bad) == -1
    assert b"multiples of 8" in capi.lib().qt_last_error()


def test_module_mirror_surface_and_no_cpu_fallback():
    from qtcnn_b200 import models as M
    m = M.QuadtreeCNN(num_classes=8)
    sd = m.state_dict()
    assert len(sd) == 252
    assert sum(p.numel() for p in m.parameters()) == 26_488_272
    for key in ("base_cnn.conv1.weight", "features_extractor.0.weight", "features_extractor.6.0.downsample.1.running_var",
                "global_processor.0.1.bn2.num_batches_tracked", "quadrant_processor.0.bias", "numerical_mlp.3.weight",
                "classifier.0.weight", "classifier.3.bias", "base_cnn.fc.weight"):
        assert key in sd, key
    assert sd["classifier.0.weight"].shape == (2688, 5376)
    assert sd["features_extractor.0.weight"].data_ptr() == sd["base_cnn.conv1.weight"].data_ptr()
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        m(torch.randn(1, 3, 224, 224), torch.randn(1, 47))
    with pytest.raises(ValueError):
        M.QuadtreeCNN(num_classes=8, mode="bogus")
    frozen = M.get_model_resnet(8, "cpu", mode="image_only", print_num_params=False)
    assert not any(p.requires_grad for p in frozen.base_cnn.parameters())
    assert frozen.classifier[0].in_features == 5120


def test_cnn_lstm_module_surface():
    """CnnLstm mirrors cnn+lstm/models.py:14-57: same state_dict keys as the reference (= the oracle's parameter names,
    which tests/golden/cnn_lstm_train_b2.json loaded into the unmodified reference), 12.68 M parameters of which
    1.50 M train, CPU tensors are refused (no fallback path)."""
    import pytest
    import torch
    from oracle import quadtree_oracle as O
    from qtcnn_b200 import models as M
    m = M.CnnLstm(8, sequence_length=3)
    p = O.make_params("cnn_lstm", 8, seed=1)
    assert set(m.state_dict()) == set(p)
    assert all(tuple(m.state_dict()[k].shape) == tuple(v.shape) for k, v in p.items())
    assert sum(q.numel() for q in m.parameters() if q.requires_grad) == 1_502_472
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 3, 64, 64), torch.zeros(1, 3, 47))
    with pytest.raises(ValueError):
        M.get_model_seq("nope", 8, "cpu")


def test_ncu_summary_is_reproducible_from_the_committed_launch_list(tmp_path):
    """profiles/ncu_gemm_summary.json (read by bench.py for roofline.traffic / tensor-pipe numbers) is exactly what
    tools/ncu_summary.py derives from the committed ncu CSV of one training step."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "profiles", "r02_ncu_gemm_launches.csv")  # the capture ncu_gemm_summary.json was last regenerated from
    out = tmp_path / "summary.json"
    subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), src, str(out)], check=True, capture_output=True)
    with open(out) as f, open(os.path.join(root, "profiles", "ncu_gemm_summary.json")) as g:
        new, committed = json.load(f), json.load(g)
    assert new == committed
    assert committed["conv3x3_kernel"]["launches_per_step"] == 26 and committed["wgrad3x3_kernel"]["launches_per_step"] == 13
    assert 0 < committed["flop_weighted_tensor_active_pct"] < 100


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path: the unmodified reference model from baseline/_ref when staged,
    else the oracle port) prints one JSON line with the keys the driver reads; non-zero ranks print nothing."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-batch", "2"]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, cwd=root, env={**os.environ, "RANK": "0"}).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f32"
    staged = os.path.exists(os.path.join(root, "baseline", "_ref", "Quadtree_from scratch", "models.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    silent = subprocess.run(cmd, check=True, capture_output=True, text=True, cwd=root, env={**os.environ, "RANK": "1"}).stdout.strip()
    assert silent == ""
