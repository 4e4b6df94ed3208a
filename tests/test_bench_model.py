"""CPU tests of bench.py's byte model and of how it picks the dominant kernel (no GPU, no timing)."""
import importlib.util
import os

from conftest import ROOT

spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_algorithmic_bytes_follow_design_md():
    n = 1000.0
    rows = [n / 8 ** l for l in range(16)]
    ab = lambda k, prec="mixed", pre="mg": bench.algorithmic_bytes_per_launch(k, n, prec, rows, pre)
    # mixed precision, per unknown
    assert ab("xpay") == 20 * n and ab("spmv_dot") == 32 * n and ab("axpy2_norm") == 52 * n
    # section 4: sweep variants and the fused residual + restriction, level 1 has an eighth of the rows
    assert ab("sweep@0z") == 24 * n and ab("sweep@0") == 28 * n and ab("sweep@0p") == 28.5 * n and ab("sweep@0d") == 28 * n
    assert ab("sweep@1") == 28 * n / 8 and ab("residual_restrict@0") == 24.5 * n
    # plain CG in fp64 and all-float CG (11V + 3C + 1 with the mask folded away: V = 8 / 4)
    assert ab("xpay", "fp64", "none") == 24 * n and ab("axpy2_norm", "fp64", "none") == 48 * n and ab("spmv_dot", "fp64", "none") == 48 * n
    assert ab("xpay", "fp32") == 12 * n and ab("axpy2_norm", "fp32") == 24 * n
    # kernels outside the byte model, and the gathered levels of a z-slab run
    assert ab("build_system") is None and ab("sweep@g0") is None


def test_dominant_kernel_groups_the_sweep_variants():
    n = 1.0e6
    rows = [n / 8 ** l for l in range(16)]
    ab = lambda k: bench.algorithmic_bytes_per_launch(k, n, "mixed", rows, "mg")
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
