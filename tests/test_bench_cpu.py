"""bench.py's host-side pieces on the CPU: the three workloads produce batches in BatchGenerator's format, both arms describe
the same `config`, and the reference arm runs the full per-GPU batch."""
import json
import subprocess
import sys
import os

import numpy as np

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workload_batches_have_batchgenerator_format():
    from myolo.config import resolve
    for name, (S, B) in {"c2": (224, 2), "c3": (416, 2), "c5": (640, 1)}.items():
        cfg = bench.bench_config(B, S, name)
        c = resolve(cfg)
        assert c["G"] == S // 32 and c["R"] == c["G"] ** 2 * c["NB"]
        assert (c["NB"], c["NC"]) == {"c2": (3, 4), "c3": (5, 2), "c5": (5, 81)}[name]
        b = bench.make_host_batches(cfg, name, 1, seed=1234)[0]
        G, NB, NC, TB, M = c["G"], c["NB"], c["NC"], c["TB"], c["MAXGT"]
        assert b[0].shape == (B, S, S, 3) and b[0].dtype == np.float32 and 0.0 <= b[0].min() and b[0].max() <= 1.0
        assert b[1].shape == (B, 1, 1, 1, TB, 4) and b[2].shape == (B, G, G, NB, 5 + NC)
        assert b[3].shape == (B, M) and b[3].dtype == np.int32 and b[4].shape == (B, M, 4)
        assert b[5].shape == (B, S, S, M) and b[5].dtype == np.bool_
        n_inst = (b[3] > 0).sum(1)
        assert n_inst.min() >= 1 and b[3].max() < NC
        assert (b[2][..., 4] == 1).sum() >= B                       # YOLO targets were encoded
        b2 = bench.make_host_batches(cfg, name, 1, seed=1234)[0]
        assert all(np.array_equal(x, y) for x, y in zip(b, b2))     # seeded -> the two arms see the same batch


def test_reference_arm_line_runs_full_batch_and_shares_the_config():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--batch", "2", "--size", "64"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/sec" and line["higher_is_better"] is True
    assert line["sample_images_per_step"] == 2 == line["config"]["per_gpu_batch"]
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0

    class A:
        config, size, batch = "c2", 64, 2
    from myolo.config import resolve
    ours = bench.workload_config(A, resolve(bench.bench_config(2, 64, "c2")), 1)
    assert ours == line["config"]
