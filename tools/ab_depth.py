import sys, subprocess, json, os
for d in (1, 2, 1, 2, 4):
    env = dict(os.environ, MLDSA_DEPTH=str(d))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-verify-leg"], env=env, capture_output=True, text=True).stdout
    j = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print("depth", d, "signs/s %.0f ms %.1f" % (j["value"], j["ms_per_step"]), "rounds", j["sign"]["rounds_per_step"], {a: round(b["ms_per_step"], 1) for a, b in k.items()}, "e2e %.0f" % j["e2e"]["value"], flush=True)
