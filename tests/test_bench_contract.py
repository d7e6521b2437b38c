"""bench.py's reference arm (the only leg that runs without a GPU): one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest

import refbind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not refbind.available("glibc"), reason="oracle/_ref/libref_glibc.so not built")
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "512", "--steps", "1",
                          "--warmup", "0", "--ref-threads", "2"], capture_output=True, text=True, timeout=300, check=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "lsd_source_mpixel_per_s" and d["unit"] == "Mpixel/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
