"""The reference arm of bench.py runs without a GPU: check its JSON line against the contract the driver parses
(one line, metric/unit/config of BASELINE.json, impl, cpu_baseline, e2e with zero copy bytes)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "2048"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["metric"] == "grid_point_field_gradient_evals_per_s" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 2048) < 1e-6 * 2048
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "2048" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert isinstance(base, dict)          # BASELINE.json is present and parses (metric naming is checked above)


def test_reference_arm_torch_backend():
    """north_star: the reference's PyTorch-CPU backend is timed too (--backend torch, oracle/gempy_oracle_torch.py)."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--backend", "torch", "--steps", "1",
                        "--warmup", "0", "--cpu-sample", "1024", "--cpu-chunk", "8000000"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.strip().startswith("{")][0])
    assert d["impl"] == "reference" and d["backend"] == "torch" and d["value"] > 0
    assert "PyTorch-CPU" in d["cpu_baseline"]["sample"] and d["cpu_baseline"]["kind"] == "port"


def test_torch_port_equals_numpy_port():
    import numpy as np
    sys.path.insert(0, ROOT)
    from gempy_b200 import examples as ex
    from oracle import gempy_oracle as orc, gempy_oracle_torch as ot
    for kernel in ("cubic", "exponential", "matern_5_2"):
        m = ex.synthetic_stress(n_sp_per_surface=30, n_surfaces=3, n_ori=20, resolution=(4, 4, 4))
        ii, opt, desc = m.args()
        ko = opt.kernel_options
        from gempy_b200.engine.data import AvailableKernelFunctions as K
        ko.kernel_function = K[kernel]
        st = orc.prepare_stack(ii.surface_points.sp_coords, ii.surface_points.nugget_effect_scalar,
                               desc.tensors_structure.number_of_points_per_surface, ii.orientations.dip_positions,
                               ii.orientations.dip_gradients, ii.orientations.nugget_effect_grad)
        w = orc.solve(orc.assemble_covariance(st, ko), orc.rhs(st, ko))
        x = np.random.default_rng(0).uniform(-0.4, 0.4, (200, 3))
        Z, G = orc.evaluate(st, ko, w, x, gradient=True)
        Zt, Gt = ot.evaluate(st, ko, w, x, gradient=True, chunk_elems=3000)
        assert np.abs(Z - Zt).max() < 1e-12 * max(1.0, np.abs(Z).max())
        assert np.abs(G - Gt).max() < 1e-12 * max(1.0, np.abs(G).max())
