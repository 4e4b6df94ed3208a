"""CPU tests of bench.py's byte model and of how it picks the dominant kernel (no GPU, no timing)."""
import importlib.util
import os

from conftest import ROOT

spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_algorithmic_bytes_follow_design_md():
    n = 1000.0
    ab = lambda k, prec="mixed", pre="mg": bench.algorithmic_bytes_per_launch(k, n, prec, pre)
    # mixed precision, per unknown
    assert ab("xpay") == 20 * n and ab("spmv_dot") == 32 * n and ab("axpy2_norm") == 52 * n
    # section 4: sweep variants and the fused residual + restriction, level 1 has an eighth of the rows
    assert ab("sweep@0z") == 24 * n and ab("sweep@0") == 28 * n and ab("sweep@0p") == 28.5 * n and ab("sweep@0d") == 28 * n
    assert ab("sweep@1") == 28 * n / 8 and ab("residual_restrict@0") == 24.5 * n
    # plain CG in fp64 and all-float CG (11V + 3C + 1 with the mask folded away: V = 8 / 4)
    assert ab("xpay", "fp64", "none") == 24 * n and ab("axpy2_norm", "fp64", "none") == 48 * n and ab("spmv_dot", "fp64", "none") == 48 * n
    assert ab("xpay", "fp32") == 12 * n and ab("axpy2_norm", "fp32") == 24 * n
    # s = z + beta s folded into the product (one kernel instead of two): xpay + spmv_dot minus the s round trip
    assert ab("xpay_spmv_dot") == (20 + 32 - 8) * n
    # kernels outside the byte model, and the gathered levels of a z-slab run
    assert ab("build_system") is None and ab("sweep@g0") is None


def test_advection_bytes_follow_design_md():
    # forward + record 19.3 B, backward + limiter 22 B per active face (Real = float), and the figure the document prints
    assert abs(bench.ADVECT_BYTES_PER_ACTIVE_FACE - (4 + 1 + 4 / 3 + 4 + 9 + 4 + 1 + 9 + 4 + 4)) < 1e-12
    assert round(bench.ADVECT_BYTES_PER_ACTIVE_FACE, 1) == 41.3
    with open(os.path.join(ROOT, "DESIGN.md")) as f:
        assert "= **41.3 B**" in f.read()


def test_dominant_kernel_groups_the_sweep_variants():
    n = 1.0e6
    ab = lambda k: bench.algorithmic_bytes_per_launch(k, n, "mixed", "mg")
    table = {"axpy2_norm": (4, 0.60), "sweep@0z": (5, 0.38), "sweep@0": (5, 0.42), "sweep@0p": (5, 0.49), "sweep@0d": (5, 0.45),
             "sweep@1": (20, 0.20), "build_system": (1, 0.43), "spmv_dot": (4, 0.46)}
    groups = bench.group_kernels(table, ab)
    assert set(groups) == {"axpy2_norm", "sweep@0", "sweep@1", "spmv_dot"}            # build_system has no byte model
    dom = max(groups, key=lambda k: groups[k]["ms"])
    g = groups[dom]
    assert dom == "sweep@0" and g["launches"] == 20 and abs(g["ms"] - 1.74) < 1e-12
    assert g["bytes"] == 5 * n * (24 + 28 + 28.5 + 28)                                # every launch with its own variant's bytes
    assert set(g["variants"]) == {"sweep@0z", "sweep@0", "sweep@0p", "sweep@0d"}
    assert abs(g["variants"]["sweep@0z"]["achieved"] - 24 * n / (0.38 / 5 * 1e-3) / 1e9) < 1e-9
    assert bench.base_tag("residual_restrict@3") == "residual_restrict@3" and bench.base_tag("sweep@12pd") == "sweep@12" and bench.base_tag("xpay") == "xpay"


def test_reference_arm_prints_what_it_measured(capsys):
    """--impl reference: every number comes from full runs of the (unmodified) reference inside this invocation — no iteration constants,
    no scaling to a size it did not run; the line states the grid it ran and fits its own wall clock."""
    import argparse
    import json
    import time
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "REFERENCE_ITERATIONS" not in src and "888" not in src
    args = argparse.Namespace(workload="dambreak_solid", n=512, steps=1, warmup=0, gpus=1, scaling="weak", residual=1e-4)
    bench.REFERENCE_N = 32                      # keep the CPU test short; the arm itself runs 128^3
    t0 = time.perf_counter()
    assert bench.run_reference_arm(args) == 0
    wall = time.perf_counter() - t0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"]["grid"] == [32, 32, 32] and line["config"]["sample_of"] == [512, 512, 512]
    assert line["ms_per_step"] * line["steps"] * 1e-3 <= wall
    assert line["value"] == 32.0 ** 3 / (line["ms_per_step"] * 1e-3) / 1e6
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["iterations"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_traffic_table_is_only_quoted_for_the_sources_it_was_captured_from(tmp_path, monkeypatch):
    groups = {"sweep@0": {"variants": {"sweep@0": {"launches": 2}}}}
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "profiles")
    os.makedirs(tmp_path / "shiokaze_b200" / "csrc")
    (tmp_path / "shiokaze_b200" / "csrc" / "k.cuh").write_text("v1")
    import json
    (tmp_path / "profiles" / "traffic.json").write_text(json.dumps({"_meta": {"set": "t", "workload": "w", "kernel_source_hash": bench.kernel_source_hash()},
                                                                    "sweep@0": {"bytes_per_row": 29.0}}))
    v, why = bench.ncu_traffic(groups, "sweep@0", 100.0)
    assert v == 2900.0
    (tmp_path / "shiokaze_b200" / "csrc" / "k.cuh").write_text("v2")
    v, why = bench.ncu_traffic(groups, "sweep@0", 100.0)
    assert v is None and "other kernel sources" in why
