"""BASELINE.json configurations C2, C3 and C5 at their NAMED sizes, through the
same workload functions `bench.py --config` times (bench_workloads.py), each verified against the
oracle on spot pixels:

  C2  GrowingDisk 51x51, nv=101: path A vs oracle, path B on the pixelated twin vs oracle, A-vs-B
      median difference (the reference's own check, tests/test_all.py:103-115)
  C3  Mixture(GrowingDisk+Stream) 101x101: get_p for x, v_x, tz_x, t, z (mass and light weighted) and
      mean/variance/skewness/kurtosis('v_x') against plain float64 sums over the 5.2 GB array
  C5  evaluate_ybar + ShotNoise at snr 50/100/200 at 101x101; float32-vs-float64 <= 1e-5
"""
import pytest

pytestmark = pytest.mark.gpu


def _free_gb():
    import torch
    return torch.cuda.mem_get_info()[0] / 2 ** 30


def test_c2_growing_disk_51(gpu_ok):
    import bench_workloads as W
    r = W.run_c2(steps=2, warmup=1, e2e_steps=1)
    v = r["verify"]
    assert v["path_A_vs_oracle"] < 1e-10 and v["path_B_vs_oracle"] < 1e-10, v
    assert v["median_rel_diff_A_vs_B"] < 5e-4 and v["class_api_vs_device"] < 1e-12, v
    assert v["ok"] and r["config"]["n_fft"] == 4907 and r["voxels"] == 4907 * 51 * 51


def test_c3_mixture_101_marginals_and_moments(gpu_ok):
    if _free_gb() < 30:
        pytest.skip("needs ~20 GB of free device memory")
    import bench_workloads as W
    r = W.run_c3(steps=1, warmup=1, e2e_steps=1)
    v = r["verify"]
    errs = v["errors"]
    for dist in W.C3_DISTS:
        for lw in (0, 1):
            assert errs[f"{dist}|lw={lw}"] < 1e-10, (dist, lw, errs)
    assert errs["get_mean"] < 1e-10 and errs["get_variance"] < 1e-9, errs
    assert errs["get_skewness"] < 1e-7 and errs["get_kurtosis"] < 1e-7, errs
    assert errs["ybar"] < 1e-10 and v["median_rel_diff_mixture_vs_pixelated"] < 5e-4, v
    assert v["ok"]


def test_c5_noise_and_float32_101(gpu_ok):
    if _free_gb() < 30:
        pytest.skip("needs ~20 GB of free device memory")
    import bench_workloads as W
    r = W.run_c5(steps=1, warmup=1, e2e_steps=1)
    v = r["verify"]
    assert v["float32_vs_float64"] < 1e-5 and v["float32_vs_float64"] > 1e-12, v
    for snr in (50, 100, 200):
        assert v["errors"][f"sigma_snr{snr}"] < 1e-13, v
    assert v["ok"], v
