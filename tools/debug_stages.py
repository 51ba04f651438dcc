import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1] if len(sys.argv) > 1 else "convT_fwd"
for stage in [int(a) for a in sys.argv[2:]] or (1, 2, 3, 4, 0):
    env = dict(os.environ, B200SR_DEBUG_STAGE=str(stage))
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_checks.py"), "--one", name],
                           capture_output=True, text=True, timeout=120, env=env)
        tail = (r.stdout + r.stderr).strip().splitlines()
        msg = [l for l in tail if "RESULT" in l or "Error" in l or "error" in l or "b200sr:" in l][:3]
        print(f"stage {stage}: rc={r.returncode} {msg}", flush=True)
    except subprocess.TimeoutExpired:
        print(f"stage {stage}: TIMEOUT", flush=True)
