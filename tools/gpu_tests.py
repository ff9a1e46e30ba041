#!/usr/bin/env python
"""Run GPU tests so that one faulting kernel cannot hide the rest: run the selection in one process, then re-run
every failed test id in its own process (a CUDA fault poisons the context of the process it happened in)."""
import re
import subprocess
import sys
import os

def run(args, log, timeout=900):
    with open(log, "w") as f:
        try:
            p = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-rf", "--tb=short", "-p", "no:cacheprovider", *args],
                               stdout=f, stderr=subprocess.STDOUT, timeout=timeout)
            return p.returncode
        except subprocess.TimeoutExpired:
            f.write("\nTIMEOUT\n")
            return 124

def main():
    out = os.environ.get("GPU_TEST_OUT", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    sel = sys.argv[1:] or ["tests"]
    tag = re.sub(r"[^A-Za-z0-9]+", "_", "_".join(sel))[:60]
    log = os.path.join(out, f"pytest_{tag}.log")
    rc = run(sel, log)
    text = open(log).read()
    print(text[-3000:])
    failed = re.findall(r"^FAILED (\S+)", text, flags=re.M)
    if rc != 0 and failed:
        print(f"== re-running {len(failed)} failed tests in isolation")
        still = []
        for i, t in enumerate(failed[:40]):
            l = os.path.join(out, f"iso_{tag}_{i}.log")
            r = run([t], l, timeout=180)
            tail = open(l).read()[-1500:]
            print(f"-- {t}: rc={r}")
            if r != 0:
                still.append(t)
                print(tail)
        print("== still failing in isolation:", len(still))
        for t in still:
            print("   ", t)
    return rc

if __name__ == "__main__":
    sys.exit(main())
