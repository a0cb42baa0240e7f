"""Host-side logic of bench.py that needs no GPU: the module imports without loading libmmidx or torch (the reference arm must
not map the product library), the per-step summary, the exact ground truth helper, and the static contract of the two arms."""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_bench_imports_without_the_product_library():
    for mod in ("mmidx_b200", "bench"):
        sys.modules.pop(mod, None)
    had_torch = "torch" in sys.modules
    import bench  # noqa: F401
    assert "mmidx_b200" not in sys.modules
    assert had_torch or "torch" not in sys.modules  # torch is imported lazily, by the GPU arm only


def test_summarize_mean_median_percentiles():
    import bench
    s = bench.summarize([1.0, 2.0, 3.0, 10.0], 100)
    assert s["ms_per_step"] == 4.0 and s["median_ms_per_step"] == 2.5
    assert s["value"] == 100 / 4.0e-3 and s["value_at_median"] == 100 / 2.5e-3
    assert s["timed_steps"] == 4 and abs(s["timed_region_s"] - 0.016) < 1e-12
    p = s["step_ms_percentiles"]
    assert p["max"] == 10.0 and p["p10"] <= p["p90"] <= p["p99"] <= p["max"]


def test_exact_ground_truth_helper_and_recall():
    import bench
    rng = np.random.default_rng(0)
    X = rng.integers(0, 256, size=(500, 128)).astype(np.float64)
    Q = rng.integers(0, 256, size=(7, 128)).astype(np.float64)
    gt = np.asarray(bench.exact_gt_cpu(X, Q, 10))
    d2 = ((X[None, :, :] - Q[:, None, :]) ** 2).sum(-1)
    for r in range(7):
        kth = np.sort(d2[r])[9]
        assert (d2[r][gt[r]] <= kth).all() and len(set(gt[r])) == 10
    assert bench.recall_at_k(gt, gt) == 1.0
    assert bench.recall_at_k(np.roll(gt, 1, axis=0) * 0 - 1, gt) == 0.0


def test_reference_arm_never_touches_the_product():
    """bench.py --impl reference times the CPU restatement only: run_reference must not import the product package or load
    libmmidx (the driver checks which .so files each arm maps)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_reference")
    body = ast.get_source_segment(src, fn)
    for banned in ("mmidx_b200", "libmmidx", "_capi", "torch.cuda"):
        assert banned not in body, banned
    assert '"impl": "reference"' in body and "h2d_bytes_per_step" in body and "cpu_baseline" in body


def test_gpu_arm_line_has_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"', '"scaling"',
                '"vs_baseline"', '"dtype"', '"data"', '"config"', '"roofline"', '"cpu_baseline"', '"e2e"', '"clocks"', '"gpu_launches"'):
        assert key in src, key
