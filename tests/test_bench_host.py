"""Host-side pieces of bench.py that can be checked without a GPU."""
import importlib.util
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _bench():
    spec = importlib.util.spec_from_file_location("adb_bench", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_watchdog_fires_unless_cancelled():
    b = _bench()
    fired = []
    w = b.Watchdog(0.05, lambda: fired.append(1)).start()
    time.sleep(0.3)
    assert fired == [1]
    w2 = b.Watchdog(0.2, lambda: fired.append(2)).start()
    w2.cancel()
    time.sleep(0.4)
    assert fired == [1]


def test_bench_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    # nothing that looks like a result line may have been printed
    assert not any(l.startswith("{") and "metric" in json.dumps(l) for l in r.stdout.splitlines())
