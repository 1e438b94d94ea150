"""CPU tests of bench.py's contract with the driver: the product arm refuses to run without a CUDA device (no CPU
fallback), the reference arm (the oracle port of the reference step on the host cores) prints ONE JSON line carrying the
keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=timeout)


def test_product_arm_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
    assert not r.stdout.strip(), "no bench line may be printed by a run that measured nothing"


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"] == "train_samples_per_sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["vs_baseline"] is None  # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == pytest.approx(d["value"])
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d.get("gpu_launches", 0) == 0


def test_committed_bench_lines_and_traffic_table():
    """The bench lines committed under profiles/ carry the keys of the contract (metric, value, e2e, roofline with its
    provenance, cpu_baseline, clocks, gpu_launches, parity_check), and the traffic table names the commit of its ncu capture
    and the hash of every kernel source it describes."""
    import glob
    import hashlib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    table = json.load(open(os.path.join(root, "profiles", "traffic_r02.json")))
    assert table.get("_commit")
    for name, row in table.items():
        if name.startswith("_"):
            continue
        assert row["traffic"] > 0 and row["algorithmic"] > 0 and row["traffic"] < 2 * row["algorithmic"], name
        src = os.path.join(root, row["source_file"])
        assert os.path.exists(src) and len(row["source_sha1"]) == 40, name
        hashlib.sha1(open(src, "rb").read()).hexdigest()  # (a changed source is flagged by bench.py, not an error)
    files = sorted(glob.glob(os.path.join(root, "profiles", "bench_r02", "bench_n*_default.json")))
    assert len(files) >= 4
    for fn in files:
        line = json.loads(open(fn).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                    "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "parity_check"):
            assert key in line, (fn, key)
        assert line["metric"] == "train_samples_per_sec" or line["metric"], fn
        assert line["gpu_launches"] > 0 and line["parity_check"]["ok"], fn
        r = line["roofline"]
        assert r["bound"] in ("hbm", "tensor") and 0 < r["frac"] < 1.5 and r["peak"] > 0, fn
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0, fn
        if line["n_gpus"] == 1:
            assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["value"] > 0, fn
        assert "c3_bf16" in line and line["c3_bf16"]["weak"]["value"] > 0, fn


@pytest.mark.parametrize("cfg,params", [("c1", 141_414_732), ("c3", 141_536_716), ("c5", 154_370_508)])
def test_parameter_counts_of_the_benchmarked_models(cfg, params):
    """``config.params`` of the bench line: SURVEY.md section 8 (probe of the unmodified reference: 141 414 732 for the
    default model, 141 536 716 with the NWP + PV-history branches, 154 370 508 for the deep variant) == bench.n_params ==
    the oracle model built on the meta device (no allocation)."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import conv3d_oracle as O

    kw = bench.CONFIGS[cfg]["model"]
    assert bench.n_params(kw) == params
    with torch.device("meta"):
        m = O.OracleModel(**kw)
    assert sum(p.numel() for p in m.parameters()) == params


def test_reference_arm_under_torchrun_prints_once():
    """The driver launches both arms the same way (torchrun for N > 1): rank 0 alone times the CPU reference and prints
    the line, the other ranks exit 0 without work and without output."""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "1"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
