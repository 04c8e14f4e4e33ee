import sys, subprocess, json, os
for cfg in (sys.argv[1:] or ["1", "2"]):   # depth[,pair[,kshare[,stage]]]
    f = cfg.split(",")
    d = int(f[0])
    env = dict(os.environ, MLDSA_DEPTH=f[0], MLDSA_PAIR=f[1] if len(f) > 1 else "1", MLDSA_KSHARE=f[2] if len(f) > 2 else "2", MLDSA_STAGE=f[3] if len(f) > 3 else "2")
    out = subprocess.run([sys.executable, "bench.py", "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-verify-leg"], env=env, capture_output=True, text=True).stdout
    j = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    k = j["kernels"]
    print("cfg", cfg, "signs/s %.0f ms %.1f" % (j["value"], j["ms_per_step"]), "rounds", j["sign"]["rounds_per_step"], {a: round(b["ms_per_step"], 1) for a, b in k.items()}, "e2e %.0f" % j["e2e"]["value"], flush=True)
