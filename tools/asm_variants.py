import os, sys, subprocess
for v in (0, 32):      # 0 = two-pass kernel (default), 32 = thread-per-frame walk
    env = dict(os.environ, VINSAT_ASM_VARIANT=str(v))
    out = subprocess.run([sys.executable, "tools/quick_perf.py", "1024", "1000", "10"], env=env, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("obs_assemble")]
    print("variant", v, line)
