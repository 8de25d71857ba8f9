#!/usr/bin/env python
"""dev tool: A/B the attention kernel across library builds (build/ab/librf_b200_*.so): correctness
(the attention cases of tests/test_gpu_ops.py) and isolated TFLOP/s at the two headline geometries."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [os.path.join(ROOT, "reflectionflow_b200", "librf_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "build", "ab", "*.so")))
code = ("import sys; sys.path.insert(0, %r); import tools.bench_ops as B; "
        "[B.bench_attention(n, 24, 30) for n in (4608, 5632, 4608, 5632)]; "
        "[B.bench_ln(r) for r in (4608, 5632, 4608, 5632)]" % ROOT)
for lib in libs:
    env = dict(os.environ, RF_B200_LIB=lib)
    print("=====", os.path.basename(lib), flush=True)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_ops.py"), "-q", "-k", "attention or ln_modulate",
                        "-x"], env=env, capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=ROOT)
    print(r.stdout.strip() or r.stderr[-500:], flush=True)
