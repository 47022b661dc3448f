"""bench.py's JSON line on the CPU: the device is replaced by the oracle-backed stand-in (tests/oracle_step.py) and the CUDA
calls of torch by no-ops, at a toy size, so that the code path of `python bench.py` and the keys of the measurement contract
(metric, value, e2e, roofline, cpu_baseline, clocks, gpu_launches, checks, config.workload ...) are exercised where no GPU
exists.  The numbers mean nothing here; the real run is the driver's on a B200."""
import contextlib
import io
import json
import sys

import numpy as np
import pytest

import bench
import piclas_b200.particle_step as ps
from oracle_step import OracleStep


class _Dev(OracleStep):
    def PhaseTiming(self):
        return np.array([1.0, 0.1, 2.0, 0.5])

    def LastTiming(self):
        return 1.0, 5

    def UploadParticles(self, PS, spec, elem, append=False, **kw):
        if append:
            PS, spec, elem = np.concatenate([self.PS, PS]), np.concatenate([self.spec, spec]), np.concatenate([self.elem, elem])
        super().UploadParticles(PS, spec, elem, **kw)


class _Clocks:
    def __init__(self, *a):
        pass

    def start(self):
        pass

    def stop(self):
        return {"sm_mhz": 0.0, "sm_max_mhz": 0.0, "reasons": []}


@pytest.mark.parametrize("variant", ["tria_cvwm", "ref_sf"])
def test_bench_line_carries_the_contract_keys(monkeypatch, variant):
    import torch
    monkeypatch.setattr(ps, "ParticleStep", _Dev)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(bench, "ClockSampler", _Clocks)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--nelem", "6", "--particles", "2e4", "--steps", "2", "--warmup", "1",
                                      "--e2e-steps", "1", "--cpu-particles", "2e4", "--cpu-steps", "1", "--variant", variant])
    args = bench.parse()
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        bench.run_b200(args)
    lines = [ln for ln in out.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1                                             # ONE JSON line
    line = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "checks"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1 and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in line["e2e"], k
    # E in; charge density (HDG input) + the whole PartSource (copy stream) out; the in-line variant moves PartSource only
    assert line["e2e"]["h2d_bytes_per_step"] == 6 ** 3 * 64 * 3 * 8 and line["e2e"]["d2h_bytes_per_step"] == 6 ** 3 * 64 * 5 * 8
    assert line["e2e"]["serial"]["d2h_bytes_per_step"] == 6 ** 3 * 64 * 4 * 8 and line["e2e"]["charge_only"]["d2h_bytes_per_step"] == 6 ** 3 * 64 * 8
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in line["roofline"], k
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["unit"] == "GB/s"
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in line["cpu_baseline"], k
    assert line["cpu_baseline"]["kind"] == "port" and line["gpu_launches"] > 0
    assert line["checks"]["ok"] and line["checks"]["particles"] == 20000
    assert line["config"]["tracking"] == ("triatracking" if variant == "tria_cvwm" else "refmapping")
