"""bench.py's contract pieces that do not need a GPU: the reference arm runs on the host cores and prints the same
`config` object as the b200 arm would (the driver compares them), for every --config."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_shared_config():
    sys.path.insert(0, ROOT)
    import bench

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "evals/s" and line["higher_is_better"]
    assert line["e2e"] == {"value": line["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1

    class A:
        pass

    for cfg, c in bench.CONFIGS.items():
        a = A()
        a.config, a.batch, a.channels, a.blocks, a.batches_per_step, a.slots = cfg, c["batch"], c["channels"], c["blocks"], c["batches_per_step"], 4
        conf = bench.make_config(a)
        assert conf["batch"] == c["batch"] and conf["batches_per_step"] == c["batches_per_step"] and "workload" in conf
        if cfg == 2:
            assert conf == line["config"]          # the two arms of one --config print the same object
            assert line["metric"] == bench.metric_name(256) == "nn_leaf_evals_per_sec_batch256"
