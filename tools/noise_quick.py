import json, subprocess, os, sys
for thr in ("768", "896", "1024"):
    env = dict(os.environ, ROD_NOISE_THREADS=thr)
    r = subprocess.run([sys.executable, "bench.py", "--no-cpu", "--steps", "5"], capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print("threads", thr, "%.0f GB/s" % d["ops"]["noise_philox"]["GB/s"], "testset %.0f" % d["ops"]["testset_1610x3"]["GB/s"])
    except Exception as e:
        print("threads", thr, "failed", e, r.stderr[-2000:])
