"""bench.py's output contract: the reference arm (runs here, CPU only) and the committed product-arm
lines (measured on B200, profiles/) carry every key the driver and the judge read."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "spmv_effective_hbm_gbs" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["sample"] and cb["value"] == d["value"] > 0
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_committed_product_lines_have_every_contract_key():
    for name, n in (("bench_r1_final_n1.json", 1), ("bench_r1_final_n8.json", 8)):
        text = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1]
        d = json.loads(text)
        assert (BASE_KEYS | {"clocks", "roofline"}) <= set(d), name
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["data"] == "synthetic" and d["gpu_launches"] == d["steps"] > 0
        assert d["warmup"] >= 3 and d["value"] > 0 and d["ms_per_step"] > 0
        assert "workload" in d["config"] and "l2" in d["config"]
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        c = d["clocks"]
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
        if n == 1:
            cb = d["cpu_baseline"]
            assert cb["kind"] == "reference" and cb["cores"] == 1 and cb["value"] > 0 and cb["sample"]
            assert r["traffic"] is not None


def test_round2_lines_carry_the_sharded_path_and_its_parity():
    """Round 2: the PageRank / R-MAT figures are top-level scalars (also inside `config`, which the driver keeps
    whole) and every line carries a `parity` block that was green when the line was printed."""
    ref_cfg_keys = {"workload", "rows", "nnz", "bytes_per_step", "l2", "parallelism"}
    for name, n in (("bench_r2_n1.json", 1), ("bench_r2_n2.json", 2), ("bench_r2_n8.json", 8), ("bench_r2_n1_gated.json", 1),
                    ("bench_r2_n2_gated.json", 2), ("bench_r2_n4_gated.json", 4), ("bench_r2_n8_gated.json", 8)):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert (BASE_KEYS | {"clocks", "roofline", "parity"}) <= set(d), name
        assert d["n_gpus"] == n and d["parity"]["ok"] is True
        for key in ("pagerank_iters_per_s", "pagerank_ms_per_iter", "pagerank_exchange", "rmat24_merge_frac"):
            assert key in d and key in d["config"], (name, key)
        assert ref_cfg_keys <= set(d["config"])
        pr20 = d["parity"]["pagerank_scale20"]
        assert pr20["ok"] and pr20["l1_vs_oracle_f64_max_over_transports"] <= 1e-6 and pr20["transports_bit_identical"]
        if n > 1:
            assert set(pr20["transports"]) == {"multicast", "p2p", "nccl"}
            assert d["parity"]["pagerank_scale26"]["ranks_bit_equal_across_gpus"] is True
            assert d["pagerank_speedup_vs_1gpu"] == round(d["pagerank_iters_per_s"] / d["pagerank_1gpu_iters_per_s"], 4) or \
                abs(d["pagerank_speedup_vs_1gpu"] - d["pagerank_iters_per_s"] / d["pagerank_1gpu_iters_per_s"]) < 1e-2
        else:
            assert d["cpu_baseline"]["parity_checked"] is True
            assert d["parity"]["config2_ell_bit_identical_to_cpu_reference"] is True
            assert d["config2_csr_frac"] >= 0.95  # the drop-in spmv_csr on config 2
    # the gated host-buffer call: within 5 % of two bare DMA copies (1.39 ms) on config 2, bit-identical to the device path
    g1 = json.loads(open(os.path.join(ROOT, "profiles", "bench_r2_n1_gated.json")).read().strip().splitlines()[-1])
    assert g1["e2e"]["ms_per_step"] <= 1.46 and g1["e2e"]["bit_identical_to_device_path"] is True
    assert g1["e2e"]["h2d_bytes_per_step"] == g1["e2e"]["d2h_bytes_per_step"] == 4 * g1["config"]["rows"]
    for name in ("bench_r2_n8.json", "bench_r2_n8_gated.json"):
        d8 = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert d8["pagerank_speedup_vs_1gpu"] >= 6.0  # north_star: >= 6x from 1 to 8 GPUs on R-MAT 26


def test_both_arms_describe_the_same_workload():
    """same_config: the reference arm and the product arm at N = 1 print the same workload string and the same
    base keys / values in `config` (the product arm adds its measured scalars)."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    ref = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][0])
    mine = json.loads(open(os.path.join(ROOT, "profiles", "bench_r2_n1.json")).read().strip().splitlines()[-1])
    for key in ("workload", "rows", "nnz", "bytes_per_step", "l2", "parallelism"):
        assert ref["config"][key] == mine["config"][key], key
    assert ref["metric"] == mine["metric"] and ref["unit"] == mine["unit"]
