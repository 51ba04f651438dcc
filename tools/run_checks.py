"""Run every per-op check of tests/opcheck.py in its own subprocess (a faulting kernel cannot poison the rest),
with a timeout, and write a summary to gpurun_out/checks.log. Debug helper for the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        import opcheck
        res, bad = opcheck.run(sys.argv[2])
        print("RESULT " + json.dumps({"res": res, "bad": {k: list(v) for k, v in bad.items()}}))
        return
    import opcheck
    names = sys.argv[1:] or list(opcheck.CHECKS)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    lines = []
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=180)
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if out:
                d = json.loads(out[0][7:])
                status = "FAIL" if d["bad"] else "ok"
                line = f"{status:4s} {n}: " + " ".join(f"{k}={v:.3e}" for k, v in d["res"].items())
            else:
                line = f"CRASH {n}: rc={r.returncode} " + (r.stdout[-600:] + r.stderr[-1500:]).replace("\n", " | ")
        except subprocess.TimeoutExpired:
            line = f"TIMEOUT {n}"
        print(line, flush=True)
        lines.append(line)
    with open(os.path.join(ROOT, "gpurun_out", "checks.log"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
